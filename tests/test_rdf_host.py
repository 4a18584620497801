"""Host-side logic of InterRDF / InterRDF_s on a CPU-only machine.

The kernels are replaced by an oracle-backed stand-in (tests/fake_backend.py,
installed through the engine's test hook) so that everything *around* the
kernels is exercised: API surface, block partition, frame batching, result
assembly, _conclude against golden outputs of the reference's own rdf.py.
The tests mirror pmda/test/test_rdf.py and test_rdf_s.py."""
import os

import numpy as np
import pytest
from numpy.testing import assert_almost_equal

from pmda_b200 import engine, _cabi
from pmda_b200.rdf import InterRDF, InterRDF_s
from pmda_b200.universe import Universe

from fake_backend import OracleBackend
from helpers import oracle_interrdf, oracle_interrdf_s

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture
def backend():
    be = OracleBackend(n_devices=2)
    engine._set_test_backend(be)
    engine.clear_cache()
    yield be
    engine._set_test_backend(None)
    engine.clear_cache()


@pytest.fixture(scope="module")
def u():
    rng = np.random.default_rng(42)
    L = 24.0
    pos = (rng.random((10, 300, 3)) * L).astype(np.float32)
    return Universe(pos, [L, L, L, 90, 90, 90])


@pytest.fixture(scope="module")
def sels(u):
    return u.atoms[:120], u.atoms[120:]


def test_no_gpu_means_error_not_fallback(u, sels):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    engine._set_test_backend(None)
    with pytest.raises(_cabi.RdfkError, match="no CPU fallback"):
        InterRDF(*sels).run()
    with pytest.raises(_cabi.RdfkError, match="no CPU fallback"):
        InterRDF_s(u, [[u.atoms[:1], u.atoms[1:5]]]).run()


def test_nbins(u, backend):
    s1, s2 = u.atoms[:3], u.atoms[3:]
    rdf = InterRDF(s1, s2, nbins=412).run()
    assert len(rdf.bins) == 412


def test_range(u, backend):
    s1, s2 = u.atoms[:3], u.atoms[3:]
    rmin, rmax = 1.0, 13.0
    rdf = InterRDF(s1, s2, range=(rmin, rmax)).run()
    assert rdf.edges[0] == rmin
    assert rdf.edges[-1] == rmax


def test_count_sum_and_dtype(u, sels, backend):
    rdf = InterRDF(*sels).run()
    want, vol, nf = oracle_interrdf(u, *sels)
    assert rdf.count.dtype == np.int64
    assert np.array_equal(rdf.count, want)
    assert rdf.n_frames == nf == 10
    assert rdf.volume == pytest.approx(vol, rel=1e-14)


def test_double_run(u, sels, backend):
    rdf = InterRDF(*sels).run()
    first = rdf.count.copy()
    rdf.run()
    assert np.array_equal(rdf.count, first)


@pytest.mark.parametrize("n_blocks", [1, 2, 3, 4])
@pytest.mark.parametrize("n_jobs", [1, 2])
def test_same_result(u, sels, backend, n_blocks, n_jobs):
    want, vol, nf = oracle_interrdf(u, *sels)
    prdf = InterRDF(*sels).run(n_blocks=n_blocks, n_jobs=n_jobs)
    assert np.array_equal(prdf.count, want)
    assert len(prdf._results) == n_blocks
    assert prdf._results.shape == (n_blocks, 2)
    per_block = sum(prdf._results[:, 0])
    assert np.array_equal(per_block, want)
    for b, frames in enumerate(prdf._blocks):
        c, v, _ = oracle_interrdf(u, *sels, frames=frames)
        assert np.array_equal(prdf._results[b, 0], c)
        assert float(prdf._results[b, 1]) == pytest.approx(v, rel=1e-14)


@pytest.mark.parametrize("step", [1, 2, 3])
def test_trj_len(u, sels, backend, step):
    want, vol, nf = oracle_interrdf(u, *sels, frames=range(0, 10, step))
    prdf = InterRDF(*sels).run(step=step)
    assert prdf.n_frames == nf
    assert np.array_equal(prdf.count, want)


def test_cdf(u, sels, backend):
    rdf = InterRDF(*sels).run()
    cdf = np.cumsum(rdf.count) / rdf.n_frames
    assert_almost_equal(rdf.cdf[-1], rdf.count.sum() / rdf.n_frames)
    assert_almost_equal(rdf.cdf, cdf)


def test_reduce(sels):
    rdf = InterRDF(*sels)
    res = []
    single_frame = np.empty(2, dtype=object)       # ragged: dtype=object
    single_frame[0] = np.array([1, 2])
    single_frame[1] = np.array([3])
    res = rdf._reduce(res, single_frame.copy())
    res = rdf._reduce(res, single_frame)
    assert_almost_equal(res[0], np.array([2, 4]))
    assert_almost_equal(res[1], np.array([6]))


@pytest.mark.parametrize("exclusion_block", [None, (1, 1), (3, 5)])
def test_exclusion(u, backend, exclusion_block):
    g = u.atoms[:150]
    want, _, _ = oracle_interrdf(u, g, g, exclusion_block=exclusion_block)
    rdf = InterRDF(g, g, exclusion_block=exclusion_block).run()
    assert np.array_equal(rdf.count, want)
    if exclusion_block == (1, 1):
        full, _, _ = oracle_interrdf(u, g, g)
        # exactly the N self pairs per frame disappear
        assert full.sum() - want.sum() == 10 * 150


def test_batching_does_not_change_results(u, sels, backend, monkeypatch):
    want, _, _ = oracle_interrdf(u, *sels)
    monkeypatch.setattr(engine.Config, "max_batch_frames", 3)
    before = backend.calls
    rdf = InterRDF(*sels).run(n_blocks=2)
    assert np.array_equal(rdf.count, want)
    assert backend.calls - before == 4          # 2 blocks x ceil(5 / 3)
    assert len(rdf.timing.io) == 10 and len(rdf.timing.compute) == 10


def test_mixed_box_kinds_are_split_into_runs(backend):
    rng = np.random.default_rng(1)
    pos = (rng.random((4, 60, 3)) * 20).astype(np.float32)
    dims = np.array([[20, 20, 20, 90, 90, 90], [20, 20, 20, 90, 90, 90],
                     [20, 21, 22, 80, 85, 75], [20, 20, 20, 90, 90, 90]],
                    dtype=np.float32)
    pos[2] = (0.01 + 0.98 * rng.random((60, 3))) @ \
        __import__("pmda_b200").mdamath.triclinic_vectors(dims[2], np.float64)
    u2 = Universe(pos, dims)
    want, vol, _ = oracle_interrdf(u2, u2.atoms, u2.atoms)
    rdf = InterRDF(u2.atoms, u2.atoms).run()
    assert np.array_equal(rdf.count, want)
    assert rdf.volume == pytest.approx(vol, rel=1e-14)


def test_single_frame_hook(u, sels, backend):
    rdf = InterRDF(*sels)
    rdf._prepare()
    ts = u.trajectory[4]
    res = rdf._single_frame(ts, sels)
    want, vol, _ = oracle_interrdf(u, *sels, frames=[4])
    assert np.array_equal(res[0], want) and float(res[1]) == vol


def test_conclude_against_reference_golden():
    """InterRDF._conclude / cdf: bit-equal to the reference's rdf.py run on the
    same per-block counts and volumes (tests/golden/make_golden.py)"""
    z = np.load(os.path.join(GOLDEN, "conclude_rdf.npz"))
    for c in range(int(z["n_cases"])):
        pre = "c%d_" % c
        obj = object.__new__(InterRDF)
        obj.nA, obj.nB = int(z[pre + "nA"]), int(z[pre + "nB"])
        excl = tuple(int(v) for v in z[pre + "excl"])
        obj._exclusion_block = excl if excl != (0, 0) else None
        nbins = int(z[pre + "nbins"])
        rng = tuple(float(v) for v in z[pre + "range"])
        edges = np.histogram([-1], bins=nbins, range=rng)[1]
        assert np.array_equal(edges, z[pre + "edges"])
        obj.edges = edges
        obj.bins = 0.5 * (edges[:-1] + edges[1:])
        assert np.array_equal(obj.bins, z[pre + "bins"])
        obj.n_frames = int(z[pre + "n_frames"])
        counts, vols = z[pre + "block_counts"], z[pre + "block_volumes"]
        res = np.empty((len(counts), 2), dtype=object)
        for b in range(len(counts)):
            res[b, 0] = counts[b]
            res[b, 1] = np.array(vols[b], dtype=np.float64)
        obj._results = res
        obj._conclude()
        assert np.array_equal(obj.count, z[pre + "count"])
        assert obj.volume == float(z[pre + "volume"])
        assert np.array_equal(obj.rdf, z[pre + "rdf"])
        assert np.array_equal(obj.cdf, z[pre + "cdf"])


# ---------------------------------------------------------------- InterRDF_s
@pytest.fixture(scope="module")
def ags(u):
    s1 = u.atoms[[7]]
    s2 = u.atoms[[20, 21]]
    s3 = u.atoms[[40, 41]]
    s4 = u.atoms[[60, 61]]
    return [[s1, s2], [s3, s4]]


def test_rdf_s_count_size(u, ags, backend):
    rdf_s = InterRDF_s(u, ags).run()
    assert len(rdf_s.count) == 2
    assert len(rdf_s.count[0]) == 1
    assert len(rdf_s.count[0][0]) == 2
    assert len(rdf_s.count[1]) == 2
    assert len(rdf_s.count[1][0]) == 2
    assert len(rdf_s.count[1][1]) == 2
    assert len(rdf_s.bins) == 75


@pytest.mark.parametrize("n_blocks", [1, 2, 3, 4])
@pytest.mark.parametrize("density", [True, False])
def test_rdf_s_same_result(u, ags, backend, n_blocks, density):
    counts, vol, nf = oracle_interrdf_s(u, ags)
    r = InterRDF_s(u, ags, density=density).run(n_blocks=n_blocks)
    shell = np.power(r.edges[1:], 3) - np.power(r.edges[:-1], 3)
    shell *= 4 / 3.0 * np.pi
    for s in range(2):
        assert r.count[s].dtype == np.float64
        assert np.array_equal(r.count[s], counts[s])
        if density:
            want = counts[s] / ((1 / (vol / nf)) * shell * nf)
        else:
            want = counts[s] / (shell * nf)
        np.testing.assert_allclose(r.rdf[s], want, rtol=1e-12)
    assert_almost_equal(r.cdf[0][0][0][-1],
                        r.count[0][0][0].sum() / r.n_frames)


def test_rdf_s_single_pair_equals_interrdf(u, backend):
    s1, s2 = u.atoms[[7]], u.atoms[[20]]
    one = InterRDF(s1, s2).run()
    site = InterRDF_s(u, [[s1, u.atoms[[20, 21]]]]).run()
    assert_almost_equal(one.count, site.count[0][0][0])
    assert_almost_equal(one.rdf, site.rdf[0][0][0])


def test_rdf_s_conclude_against_reference_golden():
    z = np.load(os.path.join(GOLDEN, "conclude_rdf_s.npz"))
    for c in range(int(z["n_cases"])):
        pre = "c%d_" % c
        obj = object.__new__(InterRDF_s)
        obj._density = bool(z[pre + "density"])
        obj.rdf_settings = {"bins": 75, "range": (0.0, 15.0)}
        obj.ag_shape = [[1, 2], [2, 3]]
        edges = np.histogram([-1], bins=75, range=(0.0, 15.0))[1]
        obj.edges = edges
        obj.n_frames = int(z[pre + "n_frames"])
        vols = z[pre + "block_volumes"]
        res = np.empty((len(vols), 2), dtype=object)
        for b in range(len(vols)):
            per = np.empty(2, dtype=object)
            for s in range(2):
                per[s] = z[pre + "block%d_site%d" % (b, s)]
            res[b, 0] = per
            res[b, 1] = np.array(vols[b], dtype=np.float64)
        obj._results = res
        obj._conclude()
        assert obj.volume == float(z[pre + "volume"])
        for s in range(2):
            assert np.array_equal(obj.count[s], z[pre + "count%d" % s])
            assert np.array_equal(obj.rdf[s], z[pre + "rdf%d" % s])
            assert np.array_equal(obj.cdf[s], z[pre + "cdf%d" % s])


def test_bin_edges_are_numpy_histogram_edges():
    """the fast path for the edges must give np.histogram's bits"""
    from pmda_b200.rdf import _histogram_edges
    rng = np.random.default_rng(0)
    cases = [(75, (0.0, 15.0)), (200, (0.0, 15.0)), (1, (0, 1)), (412, (1.5, 9.25)),
             (75, (np.float32(0.1), np.float64(7.3))), (7, (0, 7)), (3, [0.0, 1.0])]
    for _ in range(300):
        a, b = sorted(rng.random(2) * rng.choice([1.0, 10.0, 1e3, 1e-3]))
        if a < b:
            cases.append((int(rng.integers(1, 3000)), (float(a), float(b))))
    for bins, r in cases:
        want = np.histogram([-1], bins=bins, range=r)[1]
        got = _histogram_edges({'bins': bins, 'range': r})
        assert got.dtype == want.dtype and np.array_equal(got, want), (bins, r)
    # odd settings go through numpy itself
    for st in ({'bins': 5, 'range': (2.0, 2.0)}, {'bins': [0.0, 1.0, 4.0], 'range': (0.0, 4.0)}):
        assert np.array_equal(_histogram_edges(st), np.histogram([-1], **st)[1])
