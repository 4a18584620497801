"""Golden count vectors of the per-frame path, produced by the ORACLE
(oracle/rdf_oracle.c) on seeded synthetic inputs.

    python tests/golden/make_oracle_golden.py

The reference's own per-frame code cannot run here (MDAnalysis is absent), so
these vectors pin the oracle against its own regressions (CPU test) and give
the CUDA path a committed target (GPU test).  Inputs are regenerated from the
seed; their float32 bytes are fingerprinted in the fixture.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import c_oracle            # noqa: E402
from pmda_b200 import synthetic        # noqa: E402

CASES = {
    # name: (dims or None, n_atoms, n_frames, nbins, range, exclusion, n_g1)
    "ortho_same": ([33.0, 31.0, 35.0, 90, 90, 90], 2500, 2, 75, (0.0, 15.0), None, None),
    "ortho_two_groups_excl": ([30.0, 30.0, 30.0, 90, 90, 90], 1800, 2, 60, (0.5, 12.0), (3, 4), 600),
    "triclinic": ([42.0, 44.0, 46.0, 80.0, 75.0, 70.0], 2000, 2, 50, (0.0, 12.0), None, None),
    "triclinic_small_box": ([22.0, 24.0, 26.0, 70.0, 80.0, 65.0], 500, 1, 40, (0.0, 14.0), None, None),
    "no_box": (None, 1500, 2, 40, (0.0, 10.0), None, 500),
    "partial_pbc": ([28.0, 28.0, 0.0, 90, 90, 90], 1200, 2, 40, (0.0, 10.0), None, None),
}


def make_inputs(name):
    dims, n, F, nbins, rng, excl, n1 = CASES[name]
    seed = int(hashlib.sha256(name.encode()).hexdigest()[:8], 16)
    gen = np.random.default_rng(seed)
    if dims is None:
        pos = (gen.random((F, n, 3)) * 30.0).astype(np.float32)
    elif dims[3:] == [90, 90, 90]:
        L = np.array([d if d > 0 else 40.0 for d in dims[:3]])
        pos = (gen.random((F, n, 3)) * L).astype(np.float32)
    else:
        pos = synthetic.uniform_triclinic(F, n, dims, gen)
    i1 = np.arange(n) if n1 is None else np.arange(n1)
    i2 = np.arange(n) if n1 is None else np.arange(n1, n)
    return pos, dims, i1, i2, nbins, rng, excl


def oracle_counts(name):
    pos, dims, i1, i2, nbins, rng, excl = make_inputs(name)
    count = np.zeros(nbins, dtype=np.int64)
    for f in range(len(pos)):
        c_oracle.rdf_single_frame(pos[f][i1], pos[f][i2], dims, nbins, rng,
                                  excl, count=count, nthreads=4)
    return count, hashlib.sha256(pos.tobytes()).hexdigest()


if __name__ == "__main__":
    out = {}
    for name in CASES:
        count, digest = oracle_counts(name)
        out[name + "__count"] = count
        out[name + "__sha256"] = np.frombuffer(digest.encode(), dtype=np.uint8)
        print(name, int(count.sum()), digest[:12])
    np.savez(os.path.join(HERE, "oracle_counts.npz"), **out)
