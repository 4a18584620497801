"""InterRDF_s: the referenced atoms are gathered on the host before the upload
(pmda/rdf.py:296-301 gathers ``ag.positions`` first, too) and the epilogue of
pmda/rdf.py:311-334 (float64 tensors, sum over the blocks, rdf) runs on the
device (rdfk_sites_finish).  Both must leave the reference's results bit for
bit: compared with the host statement of the same lines and with the oracle.
CPU: stand-in backend; GPU: the kernels."""
import numpy as np
import pytest

from pmda_b200 import engine
from pmda_b200.rdf import InterRDF_s
from pmda_b200.universe import Universe

from helpers import oracle_interrdf_s


@pytest.fixture
def backend():
    from fake_backend import OracleBackend
    be = OracleBackend(n_devices=2)
    engine._set_test_backend(be)
    engine.clear_cache()
    yield be
    engine._set_test_backend(None)
    engine.clear_cache()


def _system(n_frames=7, n=900, L=21.0, seed=4, varying_box=False):
    rng = np.random.default_rng(seed)
    pos = (rng.random((n_frames, n, 3)) * L).astype(np.float32)
    dims = np.tile(np.array([L, L, L, 90, 90, 90], np.float32), (n_frames, 1))
    if varying_box:
        dims[:, :3] *= (1.0 + 0.01 * np.arange(n_frames))[:, None]
    u = Universe(pos, dims)
    pick = rng.permutation(n)
    ags = [[u.atoms[pick[[0]]], u.atoms[np.sort(pick[10:50])]],
           [u.atoms[pick[[1, 2]]], u.atoms[np.sort(pick[30:45])]],
           [u.atoms[pick[[3]]], u.atoms[pick[[60]]]]]
    return u, ags


def _host_conclude(r):
    """pmda/rdf.py:311-334 on the host, from the per-block results"""
    count = np.sum(r._results[:, 0])
    volume = np.sum(r._results[:, 1])
    vol = np.power(r.edges[1:], 3) - np.power(r.edges[:-1], 3)
    vol *= 4 / 3.0 * np.pi
    rdf = []
    for i in range(len(r.ag_shape)):
        if r._density:
            density = 1 / (volume / r.n_frames)
            rdf.append(count[i] / (density * vol * r.n_frames))
        else:
            rdf.append(count[i] / (vol * r.n_frames))
    return count, volume, rdf


def _check(u, ags, **kw):
    run_kw = {k: kw.pop(k) for k in ("n_blocks", "n_jobs", "step") if k in kw}
    r = InterRDF_s(u, ags, **kw).run(**run_kw)
    count, volume, rdf = _host_conclude(r)
    assert r.volume == volume
    assert isinstance(r.rdf, list) and len(r.rdf) == len(ags)
    for i in range(len(ags)):
        assert r.count[i].dtype == np.float64
        assert r.count[i].shape == (len(ags[i][0]), len(ags[i][1]),
                                    kw.get("nbins", 75))
        np.testing.assert_array_equal(r.count[i], count[i])
        np.testing.assert_array_equal(r.rdf[i], rdf[i])          # bit-equal
    frames = range(0, u.trajectory.n_frames, run_kw.get("step", 1))
    want, _, _ = oracle_interrdf_s(u, ags, kw.get("nbins", 75),
                                   kw.get("range", (0.0, 15.0)), frames=frames)
    for i in range(len(ags)):
        np.testing.assert_array_equal(r.count[i], want[i])
    cdf = r.cdf
    np.testing.assert_array_equal(
        cdf[0], np.cumsum(r.count[0], axis=2) / r.n_frames)
    return r


def test_device_conclude_matches_the_host_statement(backend):
    u, ags = _system()
    for nb in (1, 2, 3):
        _check(u, ags, nbins=40, range=(0.0, 9.0), n_blocks=nb)
    _check(u, ags, nbins=25, range=(1.0, 8.0), density=False, n_blocks=2)
    _check(u, ags, n_jobs=2, n_blocks=4, step=2)


def test_npt_boxes_and_the_rdf_denominator(backend):
    u, ags = _system(varying_box=True)
    r = _check(u, ags, nbins=30, range=(0.0, 8.0), n_blocks=3)
    assert r.volume == pytest.approx(
        float(sum(np.prod(d[:3].astype(np.float64))
                  for d in u.trajectory.dimensions)), rel=1e-12)


def test_only_referenced_atoms_are_staged(backend, monkeypatch):
    u, ags = _system(n=900)          # 59 of 900 atoms are referenced
    seen = []
    real = backend.rdf_sites

    def spy(xyz, F, natoms, *a, **k):
        seen.append((tuple(xyz.shape), natoms))
        return real(xyz, F, natoms, *a, **k)

    monkeypatch.setattr(backend, "rdf_sites", spy)
    _check(u, ags, nbins=20, range=(0.0, 9.0))
    used = len(np.unique(np.concatenate(
        [g.indices for pair in ags for g in pair])))
    assert used < 450
    assert seen and all(shape[1] == used and natoms == used
                        for shape, natoms in seen)
    # most of the frame referenced: whole frames are staged
    seen.clear()
    big = [[u.atoms[[0]], u.atoms[1:800]]]
    InterRDF_s(u, big, nbins=20, range=(0.0, 9.0)).run()
    assert seen and all(natoms == 900 for _, natoms in seen)


def test_single_frame_keeps_the_reference_structure(backend):
    u, ags = _system(n_frames=2)
    s = InterRDF_s(u, ags, nbins=20, range=(0.0, 9.0))
    s._prepare()
    r = s._single_frame(u.trajectory[1], None)
    assert r.shape == (2,) and r[0].shape == (3,)
    want, vol, _ = oracle_interrdf_s(u, ags, 20, (0.0, 9.0), frames=[1])
    np.testing.assert_array_equal(r[0][1], want[1])
    assert float(r[1]) == vol


@pytest.mark.gpu
def test_gpu_sites_gather_and_device_conclude():
    engine._set_test_backend(None)
    engine.clear_cache()
    rng = np.random.default_rng(31)
    n, L, F = 20000, 58.0, 24
    pos = (rng.random((F, n, 3)) * L).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    ags = []
    for ion in rng.choice(n, 24, replace=False):
        ags.append([u.atoms[[ion]],
                    u.atoms[np.sort(rng.choice(n, 128, replace=False))]])
    for cache in (0, 1 << 30):
        old = engine.Config.cache_bytes
        engine.Config.cache_bytes = cache
        try:
            for nb in (1, 3):
                _check(u, ags, nbins=75, range=(0.0, 15.0), n_blocks=nb)
            _check(u, ags, nbins=60, range=(0.5, 12.0), density=False)
        finally:
            engine.Config.cache_bytes = old
            engine.clear_cache()
    # triclinic cell, atoms outside the primary cell of some sites
    dims = [40.0, 42.0, 44.0, 70.0, 80.0, 65.0]
    from pmda_b200 import synthetic
    post = synthetic.uniform_triclinic(6, 6000, dims, rng)
    ut = Universe(post, dims)
    agt = [[ut.atoms[[i]], ut.atoms[np.sort(rng.choice(6000, 40, False))]]
           for i in rng.choice(6000, 10, replace=False)]
    _check(ut, agt, nbins=50, range=(0.0, 12.0), n_blocks=2)
