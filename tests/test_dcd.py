"""DCD ingest.  The format is restated from its
published layout (no MDAnalysis, no real DCD files here), so the tests pin the
byte layout explicitly, then check reader against writer, the host decode, and
on the GPU the raw path (file bytes -> device -> rdfk_unpack_planes)."""
import os
import struct

import numpy as np
import pytest

from pmda_b200 import engine
from pmda_b200.dcd import DCDFormatError, DCDTrajectory, write_dcd
from pmda_b200.rdf import InterRDF
from pmda_b200.universe import Universe

from fake_backend import OracleBackend


@pytest.fixture
def backend():
    be = OracleBackend(n_devices=2)
    engine._set_test_backend(be)
    engine.clear_cache()
    yield be
    engine._set_test_backend(None)
    engine.clear_cache()


def _frames(F=7, n=130, seed=5, L=21.0):
    rng = np.random.default_rng(seed)
    return (rng.random((F, n, 3)) * L).astype(np.float32)


def test_byte_layout_of_a_charmm_dcd(tmp_path):
    """header / frame records exactly as CHARMM and NAMD write them"""
    p = str(tmp_path / "a.dcd")
    pos = _frames(F=2, n=3)
    write_dcd(p, pos, [10.0, 11.0, 12.0, 90.0, 90.0, 60.0], title="hello")
    b = open(p, "rb").read()
    assert struct.unpack("<i", b[:4])[0] == 84 and b[4:8] == b"CORD"
    ic = struct.unpack("<20i", b[8:88])
    assert ic[0] == 2 and ic[8] == 0 and ic[10] == 1 and ic[11] == 0 and ic[19] == 24
    assert struct.unpack("<i", b[88:92])[0] == 84
    n_title = struct.unpack("<i", b[92:96])[0]
    assert n_title == 84 and struct.unpack("<i", b[96:100])[0] == 1
    assert b[100:105] == b"hello"
    off = 92 + 4 + 84 + 4
    assert struct.unpack("<3i", b[off:off + 12]) == (4, 3, 4)
    off += 12
    # frame 0: unit cell record A, cos(gamma), B, cos(beta), cos(alpha), C
    assert struct.unpack("<i", b[off:off + 4])[0] == 48
    cell = struct.unpack("<6d", b[off + 4:off + 52])
    assert cell[0] == 10.0 and cell[2] == 11.0 and cell[5] == 12.0
    assert cell[1] == pytest.approx(0.5) and cell[3] == 0.0 and cell[4] == 0.0
    off += 56
    for k in range(3):
        assert struct.unpack("<i", b[off:off + 4])[0] == 12
        np.testing.assert_array_equal(
            np.frombuffer(b[off + 4:off + 16], dtype="<f4"), pos[0, :, k])
        assert struct.unpack("<i", b[off + 16:off + 20])[0] == 12
        off += 20
    assert len(b) == off + 56 + 60          # one more frame, nothing else


@pytest.mark.parametrize("endian", ["<", ">"])
@pytest.mark.parametrize("cosines", [True, False])
def test_reader_returns_what_the_writer_stored(tmp_path, endian, cosines):
    p = str(tmp_path / "b.dcd")
    pos = _frames()
    dims = np.tile(np.array([21.0, 22.0, 23.0, 90.0, 90.0, 90.0]), (7, 1))
    dims[2] = [21.0, 22.0, 23.0, 80.0, 75.0, 70.0]
    dims[5] = [25.0, 25.0, 25.0, 60.0, 60.0, 90.0]
    write_dcd(p, pos, dims, cosines=cosines, endian=endian)
    t = DCDTrajectory(p)
    assert (t.n_frames, t.n_atoms) == (7, 130)
    assert t.little_endian == (endian == "<")
    assert t.dimensions.dtype == np.float32
    # right angles come back as exactly 90 (orthorhombic classification)
    np.testing.assert_array_equal(t.dimensions[0], dims[0].astype(np.float32))
    np.testing.assert_allclose(t.dimensions, dims, rtol=1e-6)
    for i in (0, 3, 6, -1):
        np.testing.assert_array_equal(t[i].positions, pos[i])
    for frames in (range(0, 7), range(1, 7, 2), [5, 0, 2]):
        b, d = t.frame_block(frames)
        assert b.flags.c_contiguous and b.dtype == np.float32
        np.testing.assert_array_equal(b, pos[list(frames)])
        np.testing.assert_array_equal(d, t.dimensions[list(frames)])
    rb = t.raw_block(range(2, 6))
    if endian == ">":
        assert rb is None
    else:
        raw, stride, offs, d = rb
        assert raw.shape == (4, stride) and stride == 56 + 3 * (4 * 130 + 8)
        for k in range(3):
            plane = np.frombuffer(raw[1].tobytes()[offs[k]:offs[k] + 520], "<f4")
            np.testing.assert_array_equal(plane, pos[3, :, k])


def test_no_unit_cell_and_stale_nset(tmp_path):
    p = str(tmp_path / "c.dcd")
    pos = _frames(F=4, n=9)
    write_dcd(p, pos, None)
    t = DCDTrajectory(p)
    assert t.dimensions is None and t.n_frames == 4
    # a writer that never updated NSET: the file size decides
    b = bytearray(open(p, "rb").read())
    b[8:12] = struct.pack("<i", 0)
    open(p, "wb").write(b)
    assert DCDTrajectory(p).n_frames == 4
    # truncated last frame is ignored
    open(p, "wb").write(b[:-10])
    assert DCDTrajectory(p).n_frames == 3


def test_rejects_what_it_cannot_read(tmp_path):
    p = str(tmp_path / "d.dcd")
    write_dcd(p, _frames(F=1, n=4), None)
    good = bytearray(open(p, "rb").read())
    for off, val, msg in ((4, b"VELD", "CORD"), (8 + 4 * 8, struct.pack("<i", 2), "fixed"),
                          (8 + 4 * 11, struct.pack("<i", 1), "4-dimensional"),
                          (0, struct.pack("<i", 85), "84")):
        bad = bytearray(good)
        bad[off:off + len(val)] = val
        open(p, "wb").write(bad)
        with pytest.raises(DCDFormatError, match=msg):
            DCDTrajectory(p)


def test_analysis_over_dcd_equals_in_memory(tmp_path, backend):
    p = str(tmp_path / "e.dcd")
    pos = _frames(F=9, n=250, L=22.0)
    dims = [22.0, 22.5, 23.0, 90.0, 90.0, 90.0]
    write_dcd(p, pos, dims)
    ud, um = Universe.from_dcd(p), Universe(pos, dims)
    assert ud.trajectory.filename == p
    for kw in (dict(), dict(n_blocks=3), dict(start=1, stop=8, step=2)):
        a = InterRDF(ud.atoms[:100], ud.atoms[100:], nbins=40,
                     range=(0.0, 10.0)).run(**kw)
        b = InterRDF(um.atoms[:100], um.atoms[100:], nbins=40,
                     range=(0.0, 10.0)).run(**kw)
        np.testing.assert_array_equal(a.count, b.count)
        np.testing.assert_array_equal(a.rdf, b.rdf)


@pytest.mark.gpu
def test_gpu_raw_path_unpacks_the_file_bytes(tmp_path):
    import torch
    from pmda_b200 import _cabi
    p = str(tmp_path / "f.dcd")
    pos = _frames(F=6, n=3001, L=40.0)
    write_dcd(p, pos, [40.0, 40.0, 40.0, 90.0, 90.0, 90.0])
    t = DCDTrajectory(p)
    raw, stride, offs, _ = t.raw_block(range(1, 5))
    dev = torch.as_tensor(np.ascontiguousarray(raw), device="cuda")
    out = torch.empty((4, 3001, 3), dtype=torch.float32, device="cuda")
    _cabi.unpack_planes(dev, stride, offs[0], offs[1], offs[2], 4, 3001, out,
                        torch.cuda.current_stream().cuda_stream)
    np.testing.assert_array_equal(out.cpu().numpy(), pos[1:5])
    with pytest.raises(_cabi.RdfkError, match="aligned"):
        _cabi.unpack_planes(dev, stride, offs[0] + 1, offs[1], offs[2], 4, 3001,
                            out, torch.cuda.current_stream().cuda_stream)


@pytest.mark.gpu
def test_gpu_analysis_over_dcd_equals_in_memory(tmp_path):
    p = str(tmp_path / "g.dcd")
    rng = np.random.default_rng(9)
    L = 40.0
    pos = (rng.random((7, 5000, 3)) * L).astype(np.float32)
    dims = np.tile([L, L, L, 90.0, 90.0, 90.0], (7, 1))
    dims[:, :3] += rng.random((7, 3))
    write_dcd(p, pos, dims)
    ud, um = Universe.from_dcd(p), Universe(pos, dims)
    old = engine.Config.cache_bytes, engine.Config.batch_bytes
    try:
        for cache, batch in ((0, 128 << 20), (1 << 30, 128 << 20),
                             (0, 3 * 5000 * 12)):      # 3 frames per batch
            engine.Config.cache_bytes, engine.Config.batch_bytes = cache, batch
            engine.clear_cache()
            a = InterRDF(ud.atoms, ud.atoms).run(n_blocks=2)
            b = InterRDF(um.atoms, um.atoms).run(n_blocks=2)
            np.testing.assert_array_equal(a.count, b.count)
            assert a.volume == b.volume
            a2 = InterRDF(ud.atoms, ud.atoms).run(start=1, step=2)
            b2 = InterRDF(um.atoms, um.atoms).run(start=1, step=2)
            np.testing.assert_array_equal(a2.count, b2.count)
    finally:
        engine.Config.cache_bytes, engine.Config.batch_bytes = old
        engine.clear_cache()
