"""ctypes binding of oracle/rdf_oracle.c (ORACLE -- test infrastructure).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module; the product package never does.
"""
import ctypes
import os

import numpy as np

from . import build as _build
from .numpy_oracle import box_volume, check_box

_LIB = None

BOX_NONE, BOX_ORTHO, BOX_TRI = 0, 1, 2

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def lib():
    global _LIB
    if _LIB is None:
        path = _build.OUT
        if not os.path.exists(path):
            path = _build.build()
        L = ctypes.CDLL(path)
        L.oracle_pair_distance.restype = ctypes.c_double
        L.oracle_pair_distance.argtypes = [_f32p, _f32p, ctypes.c_int, _f32p]
        L.oracle_triclinic_wrap.restype = None
        L.oracle_triclinic_wrap.argtypes = [_f32p, ctypes.c_int64, _f32p]
        L.oracle_distance_array.restype = None
        L.oracle_distance_array.argtypes = [
            _f32p, ctypes.c_int64, _f32p, ctypes.c_int64, ctypes.c_int, _f32p,
            _f64p]
        L.oracle_np_bin.restype = ctypes.c_int
        L.oracle_np_bin.argtypes = [ctypes.c_double, ctypes.c_double,
                                    ctypes.c_double, ctypes.c_int, _f64p]
        L.oracle_rdf_frame.restype = None
        L.oracle_rdf_frame.argtypes = [
            _f32p, ctypes.c_int64, _f32p, ctypes.c_int64, ctypes.c_int, _f32p,
            ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
            _f64p, ctypes.c_int, ctypes.c_int, _i64p, ctypes.c_int]
        L.oracle_rdf_frame_cells.restype = ctypes.c_int
        L.oracle_rdf_frame_cells.argtypes = [
            _f32p, ctypes.c_int64, _f32p, ctypes.c_int64, _f32p,
            ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
            _f64p, ctypes.c_int, ctypes.c_int, _i64p, ctypes.c_int]
        L.oracle_rdf_s_frame.restype = None
        L.oracle_rdf_s_frame.argtypes = [
            _f32p, ctypes.c_int64, _f32p, ctypes.c_int64, ctypes.c_int, _f32p,
            ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
            _f64p, _f64p]
        L.oracle_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _box_args(dimensions):
    kind, b = check_box(dimensions)
    if kind == "none":
        return BOX_NONE, np.zeros(9, dtype=np.float32)
    if kind == "ortho":
        out = np.zeros(9, dtype=np.float32)
        out[:3] = b
        return BOX_ORTHO, out
    return BOX_TRI, np.ascontiguousarray(b.ravel(), dtype=np.float32)


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def edges_of(nbins, rng):
    """the edges np.histogram builds (pmda/rdf.py:105)"""
    return np.histogram([-1], bins=nbins, range=rng)[1]


def max_threads():
    return int(lib().oracle_max_threads())


def distance_array(ref, conf, box=None):
    ref, conf = _c32(ref), _c32(conf)
    kind, b = _box_args(box)
    out = np.empty((len(ref), len(conf)), dtype=np.float64)
    lib().oracle_distance_array(ref, len(ref), conf, len(conf), kind, b, out)
    return out


def np_bin(d, nbins, rng, edges=None):
    edges = edges_of(nbins, rng) if edges is None else edges
    return int(lib().oracle_np_bin(float(d), float(rng[0]), float(rng[1]),
                                   int(nbins), edges))


def rdf_single_frame(g1, g2, dimensions, nbins, rng, exclusion_block=None,
                     count=None, nthreads=1, cells=False):
    """InterRDF._single_frame -> (count int64[nbins], volume); accumulates
    into `count` when given.  ``cells=True``: through the cell list
    (orthorhombic, fully periodic boxes; same integers, fewer pairs visited);
    falls back to the full pair matrix when the box does not allow it."""
    g1, g2 = _c32(g1), _c32(g2)
    kind, b = _box_args(dimensions)
    edges = edges_of(nbins, rng)
    if count is None:
        count = np.zeros(nbins, dtype=np.int64)
    xa, xb = (0, 0) if exclusion_block is None else exclusion_block
    vol = 0.0 if dimensions is None else float(box_volume(dimensions))
    if cells and kind == 1:
        rc = lib().oracle_rdf_frame_cells(
            g1, len(g1), g2, len(g2), b, float(rng[1]), float(rng[0]),
            float(rng[1]), int(nbins), edges, int(xa), int(xb), count,
            int(nthreads))
        if rc == 0:
            return count, vol
    lib().oracle_rdf_frame(g1, len(g1), g2, len(g2), kind, b, float(rng[1]),
                           float(rng[0]), float(rng[1]), int(nbins), edges,
                           int(xa), int(xb), count, int(nthreads))
    return count, vol


def rdf_s_single_frame(site_pairs, dimensions, nbins, rng, counts=None):
    """InterRDF_s._single_frame -> (list of float64 [n1,n2,nbins], volume)."""
    kind, b = _box_args(dimensions)
    edges = edges_of(nbins, rng)
    if counts is None:
        counts = [np.zeros((len(p1), len(p2), nbins), dtype=np.float64)
                  for p1, p2 in site_pairs]
    for (p1, p2), cnt in zip(site_pairs, counts):
        p1, p2 = _c32(p1), _c32(p2)
        lib().oracle_rdf_s_frame(p1, len(p1), p2, len(p2), kind, b,
                                 float(rng[1]), float(rng[0]), float(rng[1]),
                                 int(nbins), edges, cnt.reshape(-1))
    vol = 0.0 if dimensions is None else float(box_volume(dimensions))
    return counts, vol
