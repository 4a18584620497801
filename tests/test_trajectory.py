"""File-backed trajectory ingest: the frame blocks the
engine reads from an NpyTrajectory must be the stored frames, and analyses over
the file must equal analyses over the same frames in memory."""
import os

import numpy as np
import pytest

from pmda_b200 import engine
from pmda_b200.rdf import InterRDF, InterRDF_s
from pmda_b200.universe import NpyTrajectory, Universe

from fake_backend import OracleBackend


@pytest.fixture
def backend():
    be = OracleBackend(n_devices=2)
    engine._set_test_backend(be)
    engine.clear_cache()
    yield be
    engine._set_test_backend(None)
    engine.clear_cache()


@pytest.fixture
def files(tmp_path):
    rng = np.random.default_rng(77)
    L = 22.0
    pos = (rng.random((9, 250, 3)) * L).astype(np.float32)
    dims = np.tile(np.array([L, L, L, 90, 90, 90], np.float32), (9, 1))
    dims[:, :3] += rng.random((9, 3)).astype(np.float32)
    p, d = str(tmp_path / "traj.npy"), str(tmp_path / "dims.npy")
    NpyTrajectory.write(p, pos, d, dims)
    return p, d, pos, dims


def test_frames_and_blocks_equal_the_stored_arrays(files):
    p, d, pos, dims = files
    t = NpyTrajectory(p, d)
    assert t.n_frames == 9 and t.n_atoms == 250 and t.filename == p
    assert not t.positions.flags.writeable          # read-only mapping
    for i in (0, 4, 8, -1):
        ts = t[i]
        np.testing.assert_array_equal(ts.positions, pos[i])
        np.testing.assert_array_equal(ts.dimensions, dims[i])
    for frames in (range(0, 9), range(2, 9, 3), range(1, 2), [7, 1, 4]):
        b, bd = t.frame_block(frames)
        np.testing.assert_array_equal(b, pos[list(frames)])
        np.testing.assert_array_equal(bd, dims[list(frames)])
    # contiguous ranges are views of the mapping, not copies
    b, _ = t.frame_block(range(3, 7))
    assert b.base is not None and b.flags.c_contiguous


def test_rejects_wrong_layout(tmp_path):
    bad = str(tmp_path / "bad.npy")
    np.save(bad, np.zeros((3, 5, 3), dtype=np.float64))
    with pytest.raises(ValueError, match="float32"):
        NpyTrajectory(bad)
    np.save(bad, np.zeros((3, 5, 2), dtype=np.float32))
    with pytest.raises(ValueError, match="n_frames, n_atoms, 3"):
        NpyTrajectory(bad)


def test_analysis_over_file_equals_in_memory(files, backend):
    p, d, pos, dims = files
    uf = Universe.from_npy(p, d)
    um = Universe(pos, dims)
    for kw in (dict(), dict(n_blocks=3), dict(start=1, stop=8, step=2)):
        a = InterRDF(uf.atoms[:100], uf.atoms[100:], nbins=40,
                     range=(0.0, 10.0)).run(**kw)
        b = InterRDF(um.atoms[:100], um.atoms[100:], nbins=40,
                     range=(0.0, 10.0)).run(**kw)
        np.testing.assert_array_equal(a.count, b.count)
        np.testing.assert_array_equal(a.rdf, b.rdf)
        assert a.volume == b.volume
    ags_f = [[uf.atoms[[0]], uf.atoms[5:30]], [uf.atoms[[1, 2]], uf.atoms[40:50]]]
    ags_m = [[um.atoms[[0]], um.atoms[5:30]], [um.atoms[[1, 2]], um.atoms[40:50]]]
    sf = InterRDF_s(uf, ags_f, nbins=30, range=(0.0, 9.0)).run(n_blocks=2)
    sm = InterRDF_s(um, ags_m, nbins=30, range=(0.0, 9.0)).run(n_blocks=2)
    for x, y in zip(sf.count, sm.count):
        np.testing.assert_array_equal(x, y)


def test_small_batches_cover_every_frame_once(files, backend, monkeypatch):
    p, d, pos, dims = files
    uf = Universe.from_npy(p, d)
    base = InterRDF(uf.atoms, uf.atoms, nbins=25, range=(0.0, 8.0)).run()
    # 2 frames per batch: the pipeline must stitch the blocks together
    monkeypatch.setattr(engine.Config, "batch_bytes", 2 * 250 * 12)
    engine.clear_cache()
    small = InterRDF(uf.atoms, uf.atoms, nbins=25, range=(0.0, 8.0)).run()
    np.testing.assert_array_equal(small.count, base.count)
    assert small.n_frames == 9


@pytest.mark.gpu
def test_gpu_file_backed_equals_in_memory(tmp_path):
    rng = np.random.default_rng(78)
    L = 40.0
    pos = (rng.random((6, 5000, 3)) * L).astype(np.float32)
    p = str(tmp_path / "traj.npy")
    NpyTrajectory.write(p, pos)
    uf = Universe.from_npy(p, [L, L, L, 90, 90, 90])
    um = Universe(pos, [L, L, L, 90, 90, 90])
    for cache in (0, 1 << 30):
        engine.Config.cache_bytes = cache
        engine.clear_cache()
        a = InterRDF(uf.atoms, uf.atoms).run(n_blocks=2)
        b = InterRDF(um.atoms, um.atoms).run(n_blocks=2)
        np.testing.assert_array_equal(a.count, b.count)
        # a second run over the file (device cache hit when enabled)
        a2 = InterRDF(uf.atoms, uf.atoms).run()
        np.testing.assert_array_equal(a2.count, b.count)
    engine.Config.cache_bytes = 48 << 30
    engine.clear_cache()
