"""Host-side box helpers the hot path needs from MDAnalysis.

pmda calls ``capped_distance(..., box=u.dimensions)`` (pmda/rdf.py:123-126) and
reads ``ts.volume`` (pmda/rdf.py:135, :307).  MDAnalysis turns the six-number
``dimensions`` into either three float32 lengths or a float32 3x3 matrix on the
host, in float64 numpy; the trig is not reproducible bit-for-bit on a GPU, so
this stays host code.  These functions restate, from the published upstream
source (not available offline -- see SURVEY.md section 8c):

    MDAnalysis.lib.mdamath.triclinic_vectors / box_volume
    MDAnalysis.lib.util.check_box
"""
import numpy as np

from ._cabi import BOX_NONE, BOX_ORTHO, BOX_TRI


def triclinic_vectors(dimensions, dtype=np.float32):
    """[lx, ly, lz, alpha, beta, gamma] -> 3x3 lower-triangular box matrix."""
    d = np.asarray(dimensions, dtype=np.float64)
    lx, ly, lz, alpha, beta, gamma = (float(v) for v in d)
    valid = bool(np.all(d > 0.0)) and alpha < 180.0 and beta < 180.0 \
        and gamma < 180.0
    if not valid:
        return np.zeros((3, 3), dtype=dtype)
    if alpha == 90.0 and beta == 90.0 and gamma == 90.0:
        return np.diag(d[:3].astype(dtype))
    # exact values for right angles, libm otherwise
    ca = 0.0 if alpha == 90.0 else np.cos(np.deg2rad(alpha))
    cb = 0.0 if beta == 90.0 else np.cos(np.deg2rad(beta))
    if gamma == 90.0:
        cg, sg = 0.0, 1.0
    else:
        g = np.deg2rad(gamma)
        cg, sg = np.cos(g), np.sin(g)
    m = np.zeros((3, 3), dtype=np.float64)
    m[0, 0] = lx
    m[1, 0] = ly * cg
    m[1, 1] = ly * sg
    m[2, 0] = lz * cb
    m[2, 1] = lz * (ca - cb * cg) / sg
    m[2, 2] = np.sqrt(lz * lz - m[2, 0] ** 2 - m[2, 1] ** 2)
    if not m[2, 2] > 0.0:          # also catches NaN: invalid angle triplet
        return np.zeros((3, 3), dtype=dtype)
    return m.astype(dtype)


def box_volume(dimensions):
    """``Timestep.volume``: product of the lengths (ortho) or of the box-matrix
    diagonal (triclinic), in float64."""
    d = np.asarray(dimensions, dtype=np.float64)
    lx, ly, lz, alpha, beta, gamma = (float(v) for v in d)
    if alpha == 90.0 and beta == 90.0 and gamma == 90.0 \
            and lx > 0 and ly > 0 and lz > 0:
        return lx * ly * lz
    m = triclinic_vectors(d, dtype=np.float64)
    return m[0, 0] * m[1, 1] * m[2, 2]


def classify_box(dimensions):
    """``check_box``: -> (boxkind, float32[9]).

    ortho    : [lx, ly, lz, 0...]            (all angles exactly 90)
    tri_vecs : flattened float32 triclinic_vectors
    none     : ``dimensions is None``

    Deviation from upstream: an all-zero ``dimensions`` (MDAnalysis 1.x's way
    of saying "no box") is treated like ``None`` instead of producing the NaNs
    upstream's triclinic branch would give.
    """
    out = np.zeros(9, dtype=np.float32)
    if dimensions is None:
        return BOX_NONE, out
    b = np.asarray(dimensions, dtype=np.float32)
    if b.shape != (6,):
        raise ValueError("Invalid box information. Must be of the form "
                         "[lx, ly, lz, alpha, beta, gamma].")
    if not np.any(b):
        return BOX_NONE, out
    if np.all(b[3:] == 90.0):
        out[:3] = b[:3]
        return BOX_ORTHO, out
    out[:] = triclinic_vectors(b).ravel()
    return BOX_TRI, out
