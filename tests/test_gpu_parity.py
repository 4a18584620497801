"""GPU parity: the CUDA path (through the public API and the C ABI) against the
C oracle on the same seeded inputs.  Integer counts must be bit-exact; rdf
within 1e-6 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from pmda_b200 import synthetic, mdamath
from pmda_b200.rdf import InterRDF, InterRDF_s
from pmda_b200.universe import Universe

from helpers import oracle_interrdf, oracle_interrdf_s, reference_rdf

pytestmark = pytest.mark.gpu

RTOL_RDF = 1e-6   # tolerance stated by BASELINE.json north_star


def _check(u, g1, g2, nbins=75, rng=(0.0, 15.0), excl=None, **run_kw):
    r = InterRDF(g1, g2, nbins=nbins, range=rng, exclusion_block=excl)
    r.run(**run_kw)
    start, stop, step = r.start, r.stop, r.step
    frames = range(start, stop, step)
    count, volume, nf = oracle_interrdf(u, g1, g2, nbins, rng, excl, frames)
    assert r.count.dtype == np.int64
    np.testing.assert_array_equal(r.count, count)
    assert r.n_frames == nf
    assert r.volume == pytest.approx(volume, rel=1e-14)
    if count.sum() > 0 and volume > 0:
        want = reference_rdf(count, volume, nf, len(g1), len(g2), r.edges, excl)
        np.testing.assert_allclose(r.rdf, want, rtol=RTOL_RDF, atol=0)
    return r


def test_ortho_same_group():
    rng = np.random.default_rng(11)
    L = np.array([31.0, 33.5, 29.0])
    pos = (rng.random((3, 1500, 3)) * L).astype(np.float32)
    u = Universe(pos, list(L) + [90, 90, 90])
    r = _check(u, u.atoms, u.atoms)
    # self pairs sit in bin 0
    assert r.count[0] >= 3 * 1500


def test_ortho_two_groups_rectangular():
    rng = np.random.default_rng(12)
    L = np.array([40.0, 28.0, 35.0])
    pos = (rng.random((2, 3000, 3)) * L).astype(np.float32)
    u = Universe(pos, list(L) + [90, 90, 90])
    g1 = u.atoms[rng.permutation(3000)[:700]]
    g2 = u.atoms[700:3000]
    _check(u, g1, g2)


def test_no_box():
    rng = np.random.default_rng(13)
    pos = (rng.random((2, 1100, 3)) * 30).astype(np.float32)
    u = Universe(pos, None)
    _check(u, u.atoms[:500], u.atoms[500:])


def test_ortho_unwrapped_coordinates():
    rng = np.random.default_rng(14)
    L = np.array([25.0, 25.0, 25.0])
    pos = (rng.random((2, 1200, 3)) * L * 5 - 2 * L).astype(np.float32)
    u = Universe(pos, list(L) + [90, 90, 90])
    _check(u, u.atoms, u.atoms, rng=(0.0, 12.0))


def test_partial_box_axis_without_pbc():
    rng = np.random.default_rng(15)
    pos = (rng.random((2, 900, 3)) * 30).astype(np.float32)
    u = Universe(pos, [30.0, 30.0, 0.0, 90, 90, 90])
    _check(u, u.atoms, u.atoms)


@pytest.mark.parametrize("dims", [
    [40.0, 42.0, 44.0, 80.0, 75.0, 70.0],
    [50.0, 50.0, 50.0, 60.0, 60.0, 90.0],
])
def test_triclinic(dims):
    rng = np.random.default_rng(16)
    pos = synthetic.uniform_triclinic(2, 1300, dims, rng)
    u = Universe(pos, dims)
    _check(u, u.atoms, u.atoms)
    _check(u, u.atoms[:300], u.atoms[300:])


def test_triclinic_small_box_goes_exact():
    # rmax > half the box height: the fp32 brick reduction is not provably
    # equal to the reference's 27-image search -> whole frame on the fp64 path
    dims = [20.0, 22.0, 24.0, 70.0, 80.0, 65.0]
    rng = np.random.default_rng(17)
    pos = synthetic.uniform_triclinic(1, 400, dims, rng)
    u = Universe(pos, dims)
    _check(u, u.atoms, u.atoms)


def test_exclusion_block():
    rng = np.random.default_rng(18)
    L = 30.0
    pos = (rng.random((2, 1400, 3)) * L).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    r0 = _check(u, u.atoms, u.atoms)
    r1 = _check(u, u.atoms, u.atoms, excl=(7, 7))
    assert r1.count.sum() < r0.count.sum()
    _check(u, u.atoms[:600], u.atoms[600:], excl=(3, 4))


def test_range_and_nbins():
    rng = np.random.default_rng(19)
    L = 36.0
    pos = (rng.random((2, 1100, 3)) * L).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    r = _check(u, u.atoms, u.atoms, nbins=412, rng=(1.0, 13.0))
    assert len(r.bins) == 412
    assert r.edges[0] == 1.0 and r.edges[-1] == 13.0


def test_lattice_distances_exactly_on_edges():
    # simple cubic lattice with spacing == 5 bin widths: every neighbour shell
    # at an integer multiple of `a` sits exactly on a bin edge
    m, a = 8, 1.0
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"),
                 axis=-1).reshape(-1, 3)
    pos = (g * a).astype(np.float32)[None]
    L = m * a
    u = Universe(pos, [L, L, L, 90, 90, 90])
    _check(u, u.atoms, u.atoms, nbins=20, rng=(0.0, 4.0))
    _check(u, u.atoms, u.atoms, nbins=15, rng=(0.0, 3.0))


def test_blocks_steps_and_double_run():
    rng = np.random.default_rng(20)
    L = 28.0
    pos = (rng.random((7, 600, 3)) * L).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    base = _check(u, u.atoms, u.atoms)
    for nb in (2, 3, 4):
        r = _check(u, u.atoms, u.atoms, n_blocks=nb)
        np.testing.assert_array_equal(r.count, base.count)
        np.testing.assert_array_equal(r.rdf, base.rdf)
        assert len(r._results) == nb
    for step in (2, 3):
        _check(u, u.atoms, u.atoms, step=step)
    _check(u, u.atoms, u.atoms, start=2, stop=6)
    again = InterRDF(u.atoms, u.atoms)
    again.run()
    first = again.count.copy()
    again.run()
    np.testing.assert_array_equal(again.count, first)


def test_config1_subset_matches_oracle():
    u, kw = synthetic.config1(n_frames=4)
    r = _check(u, kw["g1"], kw["g2"], kw["nbins"], kw["range"])
    assert r.count[0] >= 4 * 3000          # the self pairs (d == 0)
    np.testing.assert_allclose(r.cdf, np.cumsum(r.count) / r.n_frames)


def test_single_frame_hook():
    rng = np.random.default_rng(21)
    L = 26.0
    pos = (rng.random((3, 500, 3)) * L).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    r = InterRDF(u.atoms, u.atoms)
    r._prepare()
    ts = u.trajectory[1]
    res = r._single_frame(ts, (u.atoms, u.atoms))
    count, volume, _ = oracle_interrdf(u, u.atoms, u.atoms, frames=[1])
    np.testing.assert_array_equal(res[0], count)
    assert float(res[1]) == volume


def test_interrdf_s():
    rng = np.random.default_rng(22)
    L = 24.0
    pos = (rng.random((6, 400, 3)) * L).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    ags = [[u.atoms[[3]], u.atoms[10:40]],
           [u.atoms[[5, 6]], u.atoms[100:117]],
           [u.atoms[200:203], u.atoms[[7, 9]]]]
    for density in (True, False):
        for nb in (1, 2, 3):
            r = InterRDF_s(u, ags, density=density).run(n_blocks=nb)
            counts, volume, nf = oracle_interrdf_s(u, ags)
            assert len(r.count) == 3
            for s in range(3):
                assert r.count[s].dtype == np.float64
                np.testing.assert_array_equal(r.count[s], counts[s])
            assert r.volume == pytest.approx(volume, rel=1e-15)
            vol = np.power(r.edges[1:], 3) - np.power(r.edges[:-1], 3)
            vol *= 4 / 3.0 * np.pi
            for s in range(3):
                if density:
                    want = counts[s] / ((1 / (volume / nf)) * vol * nf)
                else:
                    want = counts[s] / (vol * nf)
                np.testing.assert_allclose(r.rdf[s], want, rtol=RTOL_RDF)
            assert r.cdf[0].shape == (1, 30, 75)


def test_interrdf_s_triclinic():
    dims = [30.0, 31.0, 32.0, 75.0, 80.0, 85.0]
    rng = np.random.default_rng(23)
    pos = synthetic.uniform_triclinic(3, 300, dims, rng)
    u = Universe(pos, dims)
    ags = [[u.atoms[[0]], u.atoms[1:60]], [u.atoms[[100, 101]], u.atoms[150:170]]]
    r = InterRDF_s(u, ags).run()
    counts, volume, nf = oracle_interrdf_s(u, ags)
    for s in range(2):
        np.testing.assert_array_equal(r.count[s], counts[s])


def test_two_local_gpus_give_the_same_integers():
    """n_jobs = 2: frames sharded over two devices of this process"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(24)
    L = 50.0
    pos = (rng.random((9, 8000, 3)) * L).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    one = InterRDF(u.atoms, u.atoms).run(n_jobs=1, n_blocks=3)
    two = InterRDF(u.atoms, u.atoms).run(n_jobs=2, n_blocks=3)
    np.testing.assert_array_equal(two.count, one.count)
    np.testing.assert_array_equal(two.rdf, one.rdf)
    assert len(two._results) == 3
    ags = [[u.atoms[[0]], u.atoms[1:200]], [u.atoms[[5, 6]], u.atoms[300:350]]]
    s1 = InterRDF_s(u, ags).run(n_jobs=1)
    s2 = InterRDF_s(u, ags).run(n_jobs=2, n_blocks=2)
    for a, b in zip(s1.count, s2.count):
        np.testing.assert_array_equal(a, b)
