"""One InterRDF run of a BASELINE config shape (for ncu captures)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pmda_b200 import synthetic
from pmda_b200.rdf import InterRDF
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kw = {}
if len(sys.argv) > 3: kw["n_atoms"] = int(sys.argv[3])
u, a = synthetic.CONFIGS[cfg](n_frames=nf, **kw)
r = InterRDF(a["g1"], a["g2"], nbins=a["nbins"], range=a["range"])
for rep in range(2):
    r.run()
    print("kernel s", r.timing.compute.sum(), "count", r.count.sum())
