"""The parity cases that stress the pair kernel's per-warp queue, group log
and marks, run through the -DRDFK_CHECK build of the same sources
(librdfk_check.so, selected with RDFK_LIB in a child process): every bound the
kernel relies on is asserted on the device and counted in rdfk_rdf_stats;
the count must be 0 and the integers must be the oracle's.  Stands in for
compute-sanitizer's racecheck / memcheck, which is closed on this pool."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys
import numpy as np
import torch
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(tests)r)
from pmda_b200 import _cabi, engine, synthetic
from pmda_b200.rdf import InterRDF
from pmda_b200.universe import Universe
from helpers import oracle_interrdf

assert _cabi.build_has_checks(), _cabi.LIB_PATH


def stats():
    st = engine._state(torch.device("cuda", torch.cuda.current_device()))
    out = []
    for ws in st.workspaces:
        if ws is not None:
            w = ws[(-ws.data_ptr()) %% 256:]
            out.append(_cabi.rdf_stats(w, 0))
    return out


def run(u, g1, g2, nbins, rng, excl=None, oracle=True):
    want = None
    for flags in (0, 1, 2, 3, 4, 16):
        old = _cabi.set_debug_flags(flags)
        try:
            r = InterRDF(g1, g2, nbins=nbins, range=rng,
                         exclusion_block=excl).run()
        finally:
            _cabi.set_debug_flags(old)
        if want is None:
            want = oracle_interrdf(u, g1, g2, nbins, rng, excl)[0] \
                if oracle else r.count.copy()
        assert np.array_equal(r.count, want), flags
        for s in stats():
            assert s["violations"] == 0, (flags, s)


rng = np.random.default_rng(2026)
# dense liquid-like box: long queues, many drains, both table classes
L = 30.0
pos = (rng.random((3, 2500, 3)) * L).astype(np.float32)
u = Universe(pos, [L, L, L, 90, 90, 90])
run(u, u.atoms, u.atoms, 75, (0.0, 12.0))
run(u, u.atoms[:700], u.atoms[700:], 200, (1.0, 9.0), excl=None)
# many narrow bins far from the origin: candidate bins near the table's end,
# wide decision bands, many fp64 re-checks
far = (pos[:2, :1200] + np.float32(700.0)).astype(np.float32)
uf = Universe(far, [L, L, L, 90, 90, 90])
run(uf, uf.atoms, uf.atoms, 3000, (0.0, 12.0))
run(uf, uf.atoms, uf.atoms, 3000, (2.0, 12.0))
# everything in range of everything (cut-off > L/2): queues saturate
run(u, u.atoms[:1500], u.atoms[:1500], 40, (0.0, 14.9))
# points piled on a few sites: every pair of a group is a hit
sites = (rng.random((40, 3)) * L).astype(np.float32)
pile = sites[rng.integers(0, 40, (2, 2000))] + \
    (rng.random((2, 2000, 3)) * 0.01).astype(np.float32)
up = Universe(pile.astype(np.float32), [L, L, L, 90, 90, 90])
run(up, up.atoms, up.atoms, 50, (0.0, 10.0))
# triclinic, wrapped and unwrapped coordinates
dims = [28.0, 28.0, 28.0, 60.0, 60.0, 90.0]
ut = Universe(synthetic.uniform_triclinic(2, 3000, dims, rng), dims)
run(ut, ut.atoms, ut.atoms, 75, (0.0, 9.0))
# larger system, brute mode as the reference (the oracle would take minutes)
L = 60.0
big = Universe((rng.random((2, 30000, 3)) * L).astype(np.float32),
               [L, L, L, 90, 90, 90])
run(big, big.atoms, big.atoms, 200, (0.0, 15.0), oracle=False)
print("CHECK-BUILD-OK")
'''


def test_parity_cases_through_the_check_build():
    from pmda_b200.csrc import build as b
    lib = b.OUT_CHECK
    if not os.path.exists(lib):
        lib = b.build_check()           # needs nvcc; normally built by build()
    env = dict(os.environ, RDFK_LIB=lib)
    code = CHILD % {"root": ROOT, "tests": os.path.join(ROOT, "tests")}
    p = subprocess.run([sys.executable, "-c", code], env=env, cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=280)
    assert p.returncode == 0 and "CHECK-BUILD-OK" in p.stdout, p.stdout[-4000:]
