"""GPU parity at BASELINE.json's full system sizes.

The oracle is O(N1 N2) on the host, so at these sizes it checks *row
subsamples* of the pair matrix (a few hundred g1 atoms against the full g2);
the full matrix is covered by size-independent properties: the pruned/shifted
kernel must give the integers of its own brute-force mode (debug flags 3: every
group pair evaluated, per-pair minimum image only), counts of a symmetric
selection are even off the self pairs, block/step decompositions agree, and an
ideal gas has g(r) = 1 within Poisson noise."""
import os

import numpy as np
import pytest

from pmda_b200 import _cabi, synthetic
from pmda_b200.rdf import InterRDF, InterRDF_s

from helpers import oracle_interrdf, oracle_interrdf_s

pytestmark = pytest.mark.gpu


def _run(g1, g2, nbins, rng, debug=None, **kw):
    old = _cabi.set_debug_flags(0 if debug is None else int(debug))
    try:
        return InterRDF(g1, g2, nbins=nbins, range=rng).run(**kw)
    finally:
        _cabi.set_debug_flags(old)


def _rows(u, g, n, seed):
    pick = np.sort(np.random.default_rng(seed).choice(len(g), n, replace=False))
    return g[pick]


def test_config2_full_size():
    # 100k atoms, cubic 100 A, 200 bins
    u, kw = synthetic.config2(n_frames=3)
    g, nbins, rng = kw["g1"], kw["nbins"], kw["range"]
    r = _run(g, g, nbins, rng)
    brute = _run(g, g, nbins, rng, debug=3)
    np.testing.assert_array_equal(r.count, brute.count)
    # symmetric selection: every unordered pair is counted twice, plus the
    # N self pairs (d == 0) of every frame in bin 0
    assert (r.count[0] - 3 * len(g)) % 2 == 0 and np.all(r.count[1:] % 2 == 0)
    assert r.count[0] >= 3 * len(g)
    # ideal gas: g(r) = 1 within 6 sigma of the pair-count noise
    expect = r.rdf[5:]
    sigma = 1.0 / np.sqrt(np.maximum(r.count[5:] / 2.0, 1.0))
    assert np.all(np.abs(expect - 1.0) < 6.0 * sigma + 1e-3)
    # oracle on a row subsample: 300 atoms of g1 against all 100k of g2
    sub = _rows(u, g, 300, 1)
    rs = _run(sub, g, nbins, rng)
    want, vol, nf = oracle_interrdf(u, sub, g, nbins, rng)
    np.testing.assert_array_equal(rs.count, want)
    # decomposition invariance at full size
    r2 = _run(g, g, nbins, rng, n_blocks=3)
    np.testing.assert_array_equal(r2.count, r.count)


def test_config1_full_size_matches_oracle():
    # 3,000-water box, 100 frames: small enough for the full oracle
    u, kw = synthetic.config1(n_frames=100)
    r = _run(kw["g1"], kw["g2"], kw["nbins"], kw["range"], n_blocks=4)
    want, vol, nf = oracle_interrdf(u, kw["g1"], kw["g2"], kw["nbins"],
                                    kw["range"])
    np.testing.assert_array_equal(r.count, want)
    assert r.volume == pytest.approx(vol, rel=1e-14)


def test_config3_sites_full_shape_matches_oracle():
    # 64 site pairs (1 ion x 256 atoms) in a 50k-atom box
    u, kw = synthetic.config3(n_frames=6)
    r = InterRDF_s(u, kw["ags"], nbins=kw["nbins"], range=kw["range"]).run(
        n_blocks=2)
    counts, vol, nf = oracle_interrdf_s(u, kw["ags"], kw["nbins"], kw["range"])
    assert len(r.count) == 64
    for got, want in zip(r.count, counts):
        assert got.shape == (1, 256, 75)
        np.testing.assert_array_equal(got, want)


def test_config4_full_size_triclinic():
    # 250k atoms, rhombic-dodecahedron-like cell
    u, kw = synthetic.config4(n_frames=2)
    g, nbins, rng = kw["g1"], kw["nbins"], kw["range"]
    r = _run(g, g, nbins, rng)
    brute = _run(g, g, nbins, rng, debug=3)
    np.testing.assert_array_equal(r.count, brute.count)
    assert np.all(r.count[1:] % 2 == 0)
    sub = _rows(u, g, 120, 2)
    rs = _run(sub, g, nbins, rng)
    want, _, _ = oracle_interrdf(u, sub, g, nbins, rng)
    np.testing.assert_array_equal(rs.count, want)


def test_config5_full_size_1k_x_1m():
    # 1,000 x 1,000,000 atoms, slab density, 316 x 316 x 100 A
    u, kw = synthetic.config5(n_frames=2)
    g1, g2, nbins, rng = kw["g1"], kw["g2"], kw["nbins"], kw["range"]
    r = _run(g1, g2, nbins, rng)
    brute = _run(g1, g2, nbins, rng, debug=3)
    np.testing.assert_array_equal(r.count, brute.count)
    swapped = _run(g2, g1, nbins, rng)          # roles exchanged: same pairs
    np.testing.assert_array_equal(swapped.count, r.count)
    sub = g1[:100]
    rs = _run(sub, g2, nbins, rng)
    want, _, _ = oracle_interrdf(u, sub, g2, nbins, rng)
    np.testing.assert_array_equal(rs.count, want)


def test_repeated_runs_are_bit_identical():
    """work distribution is dynamic (atomic work counter, warp-level queues);
    the integers must not depend on it"""
    from pmda_b200 import engine
    u, kw = synthetic.config2(n_frames=4, n_atoms=60_000)
    g, nbins, rng = kw["g1"], kw["nbins"], kw["range"]
    first = _run(g, g, nbins, rng).count
    for rep in range(12):
        if rep % 3 == 0:
            engine.clear_cache()
        again = _run(g, g, nbins, rng, n_blocks=1 + rep % 3).count
        np.testing.assert_array_equal(again, first)
    u4, kw4 = synthetic.config4(n_frames=2, n_atoms=40_000)
    first = _run(kw4["g1"], kw4["g2"], kw4["nbins"], kw4["range"]).count
    for rep in range(6):
        again = _run(kw4["g1"], kw4["g2"], kw4["nbins"], kw4["range"]).count
        np.testing.assert_array_equal(again, first)
