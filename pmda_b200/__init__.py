"""pmda_b200 -- B200-native drop-in for pmda's InterRDF / InterRDF_s hot path.

    from pmda_b200.rdf import InterRDF, InterRDF_s
    from pmda_b200.contacts import Contacts, q1q2
    from pmda_b200.hbond_analysis import HydrogenBondAnalysis
    from pmda_b200.universe import Universe        # in-memory / .npy / DCD frames

See DESIGN.md for the path, its boundary and the kernels; INTEGRATION.md for
how pmda itself would bind the C ABI in include/rdfk.h.
"""
__version__ = "0.1.0"

__all__ = ["rdf", "contacts", "hbond_analysis", "parallel", "util",
           "universe", "dcd", "engine", "mdamath"]
