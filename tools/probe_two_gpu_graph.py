"""n_jobs=2 on a 2-GPU box: repeated small runs give the oracle's integers; prints
how many calls were answered by a graph replay (none: one thread per device per
run, and the graph keys are per thread)."""
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pmda_b200 import _cabi, engine
from pmda_b200.rdf import InterRDF
from pmda_b200.universe import Universe
from helpers import oracle_interrdf
rng = np.random.default_rng(5)
L = 34.0
pos = (rng.random((40, 1500, 3)) * L).astype(np.float32)
u = Universe(pos, [L, L, L, 90, 90, 90])
want, _, _ = oracle_interrdf(u, u.atoms, u.atoms, 60, (0.0, 11.0), None)
r0 = _cabi.graph_replay_count()
for rep in range(6):
    r = InterRDF(u.atoms, u.atoms, nbins=60, range=(0.0, 11.0)).run(n_jobs=2, n_blocks=4)
    assert np.array_equal(r.count, want), rep
print("n_jobs=2 repeated runs ok; graph replays:", _cabi.graph_replay_count() - r0, "devices", torch.cuda.device_count())
