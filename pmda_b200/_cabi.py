"""ctypes binding of librdfk.so (include/rdfk.h).

There is deliberately no fallback: if the CUDA library is missing or a call
fails, an exception is raised.  torch is used only for device memory and
streams (``tensor.data_ptr()``, ``torch.cuda.current_stream().cuda_stream``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
#: RDFK_LIB selects another build of the same sources (kernel experiments)
LIB_PATH = os.environ.get("RDFK_LIB") or os.path.join(_HERE, "csrc",
                                                      "librdfk.so")

BOX_NONE, BOX_ORTHO, BOX_TRI = 0, 1, 2

_lib = None


class RdfkError(RuntimeError):
    pass


def _declare(L):
    c_int, c_void_p, c_double = ctypes.c_int, ctypes.c_void_p, ctypes.c_double
    c_size_t, c_ll = ctypes.c_size_t, ctypes.c_longlong
    L.rdfk_version.restype = c_int
    L.rdfk_version.argtypes = []
    L.rdfk_last_error.restype = ctypes.c_char_p
    L.rdfk_last_error.argtypes = []
    L.rdfk_set_debug_flags.restype = c_int
    L.rdfk_set_debug_flags.argtypes = [c_int]
    L.rdfk_build_has_checks.restype = c_int
    L.rdfk_build_has_checks.argtypes = []
    L.rdfk_device_info.restype = c_int
    L.rdfk_device_info.argtypes = [c_int, ctypes.POINTER(c_int),
                                   ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                                   ctypes.POINTER(c_size_t)]
    L.rdfk_rdf_workspace_bytes.restype = c_size_t
    L.rdfk_rdf_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int]
    L.rdfk_rdf_hist.restype = c_int
    L.rdfk_rdf_hist.argtypes = [
        c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
        c_void_p, c_double, c_double, c_int, c_void_p, c_int, c_int, c_void_p,
        c_void_p, c_size_t, c_void_p]
    L.rdfk_rdf_sites.restype = c_int
    L.rdfk_rdf_sites.argtypes = [
        c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
        c_void_p, c_int, c_ll, c_int, c_void_p, c_double, c_double, c_int,
        c_void_p, c_void_p, c_void_p]
    L.rdfk_host_sites_finish.restype = c_int
    L.rdfk_host_sites_finish.argtypes = [c_void_p, c_int, c_ll, c_int,
                                         c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_int]
    L.rdfk_contacts.restype = c_int
    L.rdfk_contacts.argtypes = [
        c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
        c_double, c_double, c_double, c_void_p, c_ll, c_void_p]
    L.rdfk_unpack_planes.restype = c_int
    L.rdfk_unpack_planes.argtypes = [c_void_p, c_ll, c_ll, c_ll, c_ll, c_int,
                                     c_int, c_void_p, c_void_p]
    L.rdfk_hbonds_workspace_bytes.restype = c_size_t
    L.rdfk_hbonds_workspace_bytes.argtypes = [c_int, c_int, c_int]
    L.rdfk_hbonds.restype = c_int
    L.rdfk_hbonds.argtypes = [
        c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
        c_int, c_void_p, c_double, c_double, c_double, c_ll, c_void_p,
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    L.rdfk_host_hash64_rows.restype = ctypes.c_ulonglong
    L.rdfk_host_hash64_rows.argtypes = [c_void_p, c_size_t, c_size_t, c_ll,
                                        ctypes.c_ulonglong, c_int]
    L.rdfk_host_gather_rows.restype = c_int
    L.rdfk_host_gather_rows.argtypes = [c_void_p, c_size_t, c_ll, c_size_t,
                                        c_void_p, c_size_t, c_void_p, c_int]
    L.rdfk_rdf_stats.restype = c_int
    L.rdfk_rdf_stats.argtypes = [c_void_p, ctypes.POINTER(ctypes.c_ulonglong),
                                 c_void_p]
    L.rdfk_rdf_stats_async.restype = c_int
    L.rdfk_rdf_stats_async.argtypes = [c_void_p, c_void_p, c_void_p]
    L.rdfk_zero_async.restype = c_int
    L.rdfk_zero_async.argtypes = [c_void_p, c_size_t, c_void_p]
    L.rdfk_launch_count.restype = ctypes.c_ulonglong
    L.rdfk_launch_count.argtypes = []
    L.rdfk_graph_replay_count.restype = ctypes.c_ulonglong
    L.rdfk_graph_replay_count.argtypes = []
    L.rdfk_measure_fp32_peak.restype = c_int
    L.rdfk_measure_fp32_peak.argtypes = [ctypes.POINTER(c_double), c_void_p]


#: every symbol include/rdfk.h declares
EXPORTS = ("rdfk_version", "rdfk_last_error", "rdfk_set_debug_flags",
           "rdfk_build_has_checks", "rdfk_device_info",
           "rdfk_rdf_workspace_bytes", "rdfk_rdf_hist", "rdfk_rdf_sites",
           "rdfk_host_sites_finish",
           "rdfk_contacts", "rdfk_unpack_planes",
           "rdfk_hbonds_workspace_bytes", "rdfk_hbonds",
           "rdfk_host_hash64_rows", "rdfk_host_gather_rows",
           "rdfk_rdf_stats", "rdfk_rdf_stats_async", "rdfk_zero_async",
           "rdfk_launch_count", "rdfk_graph_replay_count",
           "rdfk_measure_fp32_peak")


def lib():
    """Load librdfk.so; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            # not built yet (fresh checkout): compile the same CUDA sources
            # with nvcc if the toolchain is here; never a CPU substitute
            why = "not built"
            try:
                from .csrc import build as _build
                if os.path.abspath(_build.OUT) == os.path.abspath(LIB_PATH):
                    _build.build()
            except Exception as exc:        # no nvcc, compile error, ...
                why = str(exc)
            if not os.path.exists(LIB_PATH):
                raise RdfkError(
                    "CUDA extension %s is missing (%s): run `python -m "
                    "pmda_b200.csrc.build` (there is no CPU fallback)"
                    % (LIB_PATH, why))
        L = ctypes.CDLL(LIB_PATH)
        _declare(L)
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().rdfk_last_error().decode("utf-8", "replace")
        raise RdfkError("librdfk error %d: %s" % (rc, msg))


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def device_info(device=0):
    sm, maj, mnr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    smem = ctypes.c_size_t()
    check(lib().rdfk_device_info(int(device), ctypes.byref(sm),
                                 ctypes.byref(maj), ctypes.byref(mnr),
                                 ctypes.byref(smem)))
    return {"sm_count": sm.value, "cc": (maj.value, mnr.value),
            "smem_optin": smem.value}


def rdf_workspace_bytes(F, n1, n2, nbins, same_group):
    return int(lib().rdfk_rdf_workspace_bytes(int(F), int(n1), int(n2),
                                              int(nbins), int(bool(same_group))))


def rdf_hist(xyz, F, natoms, idx1, n1, idx2, n2, same_group, boxkind, box, r0,
             r1, nbins, edges, excl, hist, workspace, stream):
    """All tensor arguments are torch CUDA tensors (or None); see rdfk.h."""
    xa, xb = (0, 0) if excl is None else (int(excl[0]), int(excl[1]))
    check(lib().rdfk_rdf_hist(
        _ptr(xyz), int(F), int(natoms), _ptr(idx1), int(n1), _ptr(idx2),
        int(n2), int(bool(same_group)), int(boxkind), _ptr(box), float(r0),
        float(r1), int(nbins), _ptr(edges), xa, xb, _ptr(hist),
        _ptr(workspace), int(workspace.numel() * workspace.element_size()),
        ctypes.c_void_p(int(stream))))


def rdf_sites(xyz, F, natoms, idx_a, idx_b, off_a, off_b, cell_off, n_sites,
              total_cells, boxkind, box, r0, r1, nbins, edges, count, stream):
    check(lib().rdfk_rdf_sites(
        _ptr(xyz), int(F), int(natoms), _ptr(idx_a), _ptr(idx_b), _ptr(off_a),
        _ptr(off_b), _ptr(cell_off), int(n_sites), int(total_cells),
        int(boxkind), _ptr(box), float(r0), float(r1), int(nbins), _ptr(edges),
        _ptr(count), ctypes.c_void_p(int(stream))))


def _np_ptr(a):
    return None if a is None else \
        ctypes.c_void_p(a.__array_interface__["data"][0])


def host_sites_finish(count, nbins, den=None, want_total=False, nthreads=1):
    """rdfk_host_sites_finish on numpy arrays: uint32/int32 ``count``
    [n_blocks, n] -> (float64 [n_blocks, n], total [n] or None, rdf [n] or
    None), fresh arrays.  Releases the GIL."""
    import numpy as np
    if count.dtype not in (np.uint32, np.int32) or count.ndim != 2 or \
            not count.flags.c_contiguous:
        raise ValueError("host_sites_finish: bad count array")
    n_blocks, n = count.shape
    out = np.empty((n_blocks, n), dtype=np.float64)
    total = np.empty(n, dtype=np.float64) if want_total else None
    rdf = None
    if den is not None:
        den = np.ascontiguousarray(den, dtype=np.float64)
        if den.shape != (nbins,):
            raise ValueError("host_sites_finish: den must have nbins entries")
        rdf = np.empty(n, dtype=np.float64)
    check(lib().rdfk_host_sites_finish(
        _np_ptr(count), int(n_blocks), int(n), int(nbins), _np_ptr(den),
        _np_ptr(out), _np_ptr(total), _np_ptr(rdf), int(nthreads)))
    return out, total, rdf


CONTACT_HARD, CONTACT_RADIUS, CONTACT_SOFT, CONTACT_DIST = 0, 1, 2, 3


def contacts(xyz, F, natoms, idx_a, idx_b, r0, n_pairs, method, radius, beta,
             lam, out, out_stride, stream):
    check(lib().rdfk_contacts(
        _ptr(xyz), int(F), int(natoms), _ptr(idx_a), _ptr(idx_b), _ptr(r0),
        int(n_pairs), int(method), float(radius), float(beta), float(lam),
        _ptr(out), int(out_stride), ctypes.c_void_p(int(stream))))


def unpack_planes(raw, frame_stride, off_x, off_y, off_z, F, natoms, xyz,
                  stream):
    check(lib().rdfk_unpack_planes(
        _ptr(raw), int(frame_stride), int(off_x), int(off_y), int(off_z),
        int(F), int(natoms), _ptr(xyz), ctypes.c_void_p(int(stream))))


def hbonds_workspace_bytes(F, n_dh, n_acc):
    return int(lib().rdfk_hbonds_workspace_bytes(int(F), int(n_dh), int(n_acc)))


def hbonds(xyz, F, natoms, donors, hydrogens, n_dh, acceptors, n_acc, boxkind,
           box, max_cutoff, min_cutoff, angle_cutoff, capacity, out_row,
           out_acc, out_dist, out_angle, total, workspace, stream):
    """rdfk_hbonds: `hydrogens` None = plain capped_distance; `min_cutoff`
    None = no lower bound; `total` is a one-element int64 device tensor."""
    check(lib().rdfk_hbonds(
        _ptr(xyz), int(F), int(natoms), _ptr(donors), _ptr(hydrogens),
        int(n_dh), _ptr(acceptors), int(n_acc), int(boxkind), _ptr(box),
        float(max_cutoff), -1.0 if min_cutoff is None else float(min_cutoff),
        float(angle_cutoff), int(capacity), _ptr(out_row), _ptr(out_acc),
        _ptr(out_dist), _ptr(out_angle), _ptr(total), _ptr(workspace),
        int(workspace.numel() * workspace.element_size()),
        ctypes.c_void_p(int(stream))))


def host_hash_rows(arr, seed=0, nthreads=1):
    """64-bit content hash of every byte of a numpy array whose rows (axis 0)
    are contiguous inside (any stride between rows).  Releases the GIL."""
    if arr.ndim == 0 or arr.size == 0:
        return int(lib().rdfk_host_hash64_rows(None, 0, 0, 0, int(seed), 1))
    n_rows = arr.shape[0]
    row = arr[0]
    if not row.flags.c_contiguous:
        raise ValueError("rows must be C-contiguous")
    row_bytes = row.nbytes
    stride = arr.strides[0] if n_rows > 1 else row_bytes
    return int(lib().rdfk_host_hash64_rows(
        ctypes.c_void_p(arr.__array_interface__["data"][0]), int(n_rows),
        int(row_bytes), int(stride), int(seed) & 0xFFFFFFFFFFFFFFFF,
        int(nthreads)))


def host_gather_rows(src, idx, dst, nthreads=1):
    """dst[f, k] = src[f, idx[k]] for float32 [F, N, 3] `src` whose frames
    are C-contiguous inside (any stride between frames), int32 `idx`,
    C-contiguous float32 [F, len(idx), 3] `dst`.  Releases the GIL."""
    import numpy as np
    if src.dtype != np.float32 or dst.dtype != np.float32 or \
            idx.dtype != np.int32 or not idx.flags.c_contiguous or \
            not dst.flags.c_contiguous or src.ndim != 3 or src.shape[2] != 3 \
            or (len(src) and not src[0].flags.c_contiguous) or \
            dst.shape != (len(src), len(idx), 3):
        raise ValueError("host_gather_rows: bad array layout")
    if len(src) == 0 or len(idx) == 0:
        return
    stride = src.strides[0] if len(src) > 1 else src[0].nbytes
    check(lib().rdfk_host_gather_rows(
        ctypes.c_void_p(src.__array_interface__["data"][0]), len(src),
        int(stride), int(src.shape[1]),
        ctypes.c_void_p(idx.__array_interface__["data"][0]), len(idx),
        ctypes.c_void_p(dst.__array_interface__["data"][0]), int(nthreads)))


def rdf_stats(workspace, stream):
    out = (ctypes.c_ulonglong * 8)()
    check(lib().rdfk_rdf_stats(_ptr(workspace), out,
                               ctypes.c_void_p(int(stream))))
    return {"slow_pairs": int(out[0]), "all_slow_frames": int(out[1]),
            "tile_items": int(out[2]), "tile_pairs": int(out[3]),
            "wrapped_atoms": int(out[4]), "wrapped_frames": int(out[5]),
            "violations": int(out[6]), "first_violation": int(out[7])}


def rdf_stats_async(workspace, out_dev, stream):
    """the 8 workspace counters -> int64 device tensor `out_dev`, async"""
    check(lib().rdfk_rdf_stats_async(_ptr(workspace), _ptr(out_dev),
                                     ctypes.c_void_p(int(stream))))


def zero_async(t, stream):
    check(lib().rdfk_zero_async(_ptr(t), int(t.numel() * t.element_size()),
                                ctypes.c_void_p(int(stream))))


def set_debug_flags(flags):
    """cross-check switches of the test suite (rdfk.h); returns the old ones"""
    return int(lib().rdfk_set_debug_flags(int(flags)))


def build_has_checks():
    return bool(lib().rdfk_build_has_checks())


def launch_count():
    return int(lib().rdfk_launch_count())


def graph_replay_count():
    return int(lib().rdfk_graph_replay_count())


def measure_fp32_peak(stream=0):
    v = ctypes.c_double()
    check(lib().rdfk_measure_fp32_peak(ctypes.byref(v),
                                       ctypes.c_void_p(int(stream))))
    return float(v.value)
