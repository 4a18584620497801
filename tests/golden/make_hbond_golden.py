"""Golden hydrogen-bond tables of the HydrogenBondAnalysis path, produced by
the ORACLE (oracle/numpy_oracle.py) on seeded synthetic water boxes.

    python tests/golden/make_hbond_golden.py

MDAnalysis is absent here, so these tables pin the oracle against its own
regressions (CPU test) and give the CUDA path a committed target (GPU test).
Inputs are regenerated from the seed; their float32 bytes are fingerprinted.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import numpy_oracle        # noqa: E402
from pmda_b200 import synthetic        # noqa: E402

CASES = {
    # name: (dims, box used to build the molecules, n_mol, n_frames, d_a, angle)
    "ortho": ([19.0, 20.0, 21.0, 90, 90, 90], None, 220, 2, 3.0, 150.0),
    "triclinic": ([20.0, 21.0, 22.0, 75.0, 80.0, 70.0], None, 220, 2, 3.0, 150.0),
    "no_box": (None, [18.0, 18.0, 18.0, 90, 90, 90], 200, 2, 3.2, 140.0),
    "wide_criteria": ([17.0, 17.0, 17.0, 90, 90, 90], None, 150, 1, 4.5, 100.0),
}


def make_inputs(name):
    dims, build, n_mol, F, d_a, ang = CASES[name]
    seed = int(hashlib.sha256(("hb" + name).encode()).hexdigest()[:8], 16)
    gen = np.random.default_rng(seed)
    pos, names = synthetic.water_box(F, n_mol, build or dims, gen)
    o = 3 * np.arange(n_mol)
    don = np.repeat(o, 2)
    hyd = (o[:, None] + np.array([1, 2])).reshape(-1)
    return pos, dims, names, don, hyd, o, d_a, ang


def oracle_rows(name):
    pos, dims, names, don, hyd, acc, d_a, ang = make_inputs(name)
    rows = np.vstack([numpy_oracle.hbonds_single_frame(
        f, pos[f], don, hyd, acc, dims, d_a, ang) for f in range(len(pos))])
    return rows, hashlib.sha256(pos.tobytes()).hexdigest()


if __name__ == "__main__":
    out = {}
    for name in sorted(CASES):
        rows, digest = oracle_rows(name)
        out[name + "__rows"] = rows
        out[name + "__sha256"] = np.frombuffer(digest.encode(), dtype=np.uint8)
        print(name, rows.shape)
    np.savez_compressed(os.path.join(HERE, "hbond_rows.npz"), **out)
