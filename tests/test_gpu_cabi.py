"""The C ABI itself on the GPU (no engine, no host classes): raw device
pointers into rdfk_rdf_hist / rdfk_rdf_sites / rdfk_contacts, results against
the oracle, status codes for bad workspaces."""
import numpy as np
import pytest
import torch

from pmda_b200 import _cabi
from oracle import c_oracle, numpy_oracle

pytestmark = pytest.mark.gpu


def _dev(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda")


def test_rdf_hist_raw_call_accumulates_and_matches_oracle():
    rng = np.random.default_rng(41)
    F, n, L, nbins, r = 3, 4000, 42.0, 90, (0.0, 13.5)
    pos = (rng.random((F, n, 3)) * L).astype(np.float32)
    idx1 = np.sort(rng.choice(n, 1500, replace=False)).astype(np.int32)
    idx2 = np.arange(n, dtype=np.int32)
    box = np.zeros((F, 9), np.float32)
    box[:, :3] = L
    edges = np.histogram([-1], bins=nbins, range=r)[1]
    xyz, i1, i2 = _dev(pos, torch.float32), _dev(idx1, torch.int32), _dev(idx2, torch.int32)
    bx, ed = _dev(box, torch.float32), _dev(edges, torch.float64)
    hist = torch.zeros(nbins, dtype=torch.int64, device="cuda")
    nbytes = _cabi.rdf_workspace_bytes(F, len(idx1), len(idx2), nbins, False)
    raw = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
    off = (-raw.data_ptr()) % 256
    ws = raw[off:off + nbytes]
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):          # counts are ADDED into hist
        _cabi.rdf_hist(xyz, F, n, i1, len(idx1), i2, len(idx2), False,
                       _cabi.BOX_ORTHO, bx, r[0], r[1], nbins, ed, None, hist,
                       ws, stream)
    torch.cuda.synchronize()
    want = np.zeros(nbins, dtype=np.int64)
    for f in range(F):
        c_oracle.rdf_single_frame(pos[f][idx1], pos[f][idx2],
                                  [L, L, L, 90, 90, 90], nbins, r, count=want,
                                  nthreads=8)
    np.testing.assert_array_equal(hist.cpu().numpy(), 2 * want)
    st = _cabi.rdf_stats(ws, stream)
    assert st["tile_items"] == F * (-(-len(idx2) // 128))   # larger group = items
    assert 0 < st["tile_pairs"] < F * len(idx1) * len(idx2)
    # too small / misaligned workspace: status code, no launch
    with pytest.raises(_cabi.RdfkError, match="workspace"):
        _cabi.rdf_hist(xyz, F, n, i1, len(idx1), i2, len(idx2), False,
                       _cabi.BOX_ORTHO, bx, r[0], r[1], nbins, ed, None, hist,
                       ws[:nbytes // 2], stream)
    with pytest.raises(_cabi.RdfkError, match="workspace"):
        _cabi.rdf_hist(xyz, F, n, i1, len(idx1), i2, len(idx2), False,
                       _cabi.BOX_ORTHO, bx, r[0], r[1], nbins, ed, None, hist,
                       raw[off + 4:off + 4 + nbytes], stream)


def test_rdf_hist_null_indices_mean_all_atoms():
    rng = np.random.default_rng(42)
    F, n, L, nbins, r = 2, 1500, 30.0, 40, (0.0, 10.0)
    pos = (rng.random((F, n, 3)) * L).astype(np.float32)
    box = np.zeros((F, 9), np.float32)
    box[:, :3] = L
    edges = np.histogram([-1], bins=nbins, range=r)[1]
    hist = torch.zeros(nbins, dtype=torch.int64, device="cuda")
    nbytes = _cabi.rdf_workspace_bytes(F, n, n, nbins, True)
    raw = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
    off = (-raw.data_ptr()) % 256
    _cabi.rdf_hist(_dev(pos, torch.float32), F, n, None, n, None, n, True,
                   _cabi.BOX_ORTHO, _dev(box, torch.float32), r[0], r[1], nbins,
                   _dev(edges, torch.float64), None, hist,
                   raw[off:off + nbytes], torch.cuda.current_stream().cuda_stream)
    want = np.zeros(nbins, dtype=np.int64)
    for f in range(F):
        c_oracle.rdf_single_frame(pos[f], pos[f], [L, L, L, 90, 90, 90], nbins,
                                  r, count=want, nthreads=8)
    np.testing.assert_array_equal(hist.cpu().numpy(), want)


def test_contacts_raw_call_all_methods():
    rng = np.random.default_rng(43)
    F, n = 5, 300
    pos = (rng.random((F, n, 3)) * 20).astype(np.float32)
    ia = rng.integers(0, 150, 700).astype(np.int32)
    ib = rng.integers(150, 300, 700).astype(np.int32)
    # bit-exact distances need the reference's summation order
    d = np.stack([np.array([numpy_oracle.distance_array_nobox(
        pos[f][ia[p]][None], pos[f][ib[p]][None])[0, 0] for p in range(700)])
        for f in range(F)])
    r0 = d[0].copy()
    xyz, a, b = _dev(pos, torch.float32), _dev(ia, torch.int32), _dev(ib, torch.int32)
    r0d = _dev(r0, torch.float64)
    stream = torch.cuda.current_stream().cuda_stream
    out = torch.zeros((F, 3), dtype=torch.float64, device="cuda")
    _cabi.contacts(xyz, F, n, a, b, r0d, 700, _cabi.CONTACT_HARD, 0, 0, 0,
                   out[:, 0:], 3, stream)
    _cabi.contacts(xyz, F, n, a, b, r0d, 700, _cabi.CONTACT_RADIUS, 7.5, 0, 0,
                   out[:, 1:], 3, stream)
    _cabi.contacts(xyz, F, n, a, b, r0d, 700, _cabi.CONTACT_SOFT, 0, 5.0, 1.8,
                   out[:, 2:], 3, stream)
    dist = torch.zeros((F, 700), dtype=torch.float64, device="cuda")
    _cabi.contacts(xyz, F, n, a, b, None, 700, _cabi.CONTACT_DIST, 0, 0, 0,
                   dist, 700, stream)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    np.testing.assert_array_equal(dist.cpu().numpy(), d)
    np.testing.assert_array_equal(o[:, 0], (d <= r0).sum(1))
    np.testing.assert_array_equal(o[:, 1], (d <= 7.5).sum(1))
    np.testing.assert_allclose(o[:, 2], (1 / (1 + np.exp(5.0 * (d - 1.8 * r0)))).sum(1),
                               rtol=1e-12)
    with pytest.raises(_cabi.RdfkError, match="out_stride"):
        _cabi.contacts(xyz, F, n, a, b, None, 700, _cabi.CONTACT_DIST, 0, 0, 0,
                       dist, 10, stream)
