"""numpy restatement of the reference hot path (ORACLE -- test infrastructure).

Literal, slow, small-N version of what ``pmda.rdf`` executes per frame:

    pairs, dist = MDAnalysis.lib.distances.capped_distance(a, b, rmax, box=dims)
    count = np.histogram(dist, bins=nbins, range=rng)[0]        # pmda/rdf.py:123-134

``np.histogram`` is the *real* numpy function the reference calls, so the
binning half of this file is the reference itself.  The distance half restates
MDAnalysis (absent from /root/reference and not installable here):
``lib/distances.py`` (check_box, _bruteforce_capped, distance_array),
``lib/mdamath.py`` (triclinic_vectors, box_volume) and the C loops of
``lib/include/calc_distances.h`` -- PARITY UNPINNED for those, see
oracle/rdf_oracle.c header.  Used to cross-check the C oracle on tiny inputs.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module.
"""
import numpy as np

FLT_EPSILON = np.finfo(np.float32).eps
FLT_MAX = float(np.finfo(np.float32).max)


# -- MDAnalysis.lib.mdamath.triclinic_vectors [recalled-upstream] -------------
def triclinic_vectors(dimensions, dtype=np.float32):
    dim = np.asarray(dimensions, dtype=np.float64)
    lx, ly, lz, alpha, beta, gamma = dim
    if not (np.all(dim > 0.0) and alpha < 180.0 and beta < 180.0
            and gamma < 180.0):
        return np.zeros((3, 3), dtype=dtype)
    if alpha == beta == gamma == 90.0:
        return np.diag(dim[:3].astype(dtype))
    m = np.zeros((3, 3), dtype=np.float64)
    m[0, 0] = lx
    cos_alpha = 0.0 if alpha == 90.0 else np.cos(np.deg2rad(alpha))
    cos_beta = 0.0 if beta == 90.0 else np.cos(np.deg2rad(beta))
    if gamma == 90.0:
        cos_gamma, sin_gamma = 0.0, 1.0
    else:
        g = np.deg2rad(gamma)
        cos_gamma, sin_gamma = np.cos(g), np.sin(g)
    m[1, 0] = ly * cos_gamma
    m[1, 1] = ly * sin_gamma
    m[2, 0] = lz * cos_beta
    m[2, 1] = lz * (cos_alpha - cos_beta * cos_gamma) / sin_gamma
    m[2, 2] = np.sqrt(lz * lz - m[2, 0] ** 2 - m[2, 1] ** 2)
    if m[2, 2] > 0.0:
        return m.astype(dtype)
    return np.zeros((3, 3), dtype=dtype)


# -- MDAnalysis.lib.mdamath.box_volume [recalled-upstream] --------------------
def box_volume(dimensions):
    dim = np.asarray(dimensions, dtype=np.float64)
    lx, ly, lz, alpha, beta, gamma = dim
    if alpha == beta == gamma == 90.0 and lx > 0 and ly > 0 and lz > 0:
        return lx * ly * lz
    tv = triclinic_vectors(dimensions, dtype=np.float64)
    return tv[0, 0] * tv[1, 1] * tv[2, 2]


# -- MDAnalysis.lib.util.check_box [recalled-upstream] ------------------------
def check_box(box):
    """-> ('none', None) | ('ortho', float32[3]) | ('tri_vecs', float32[3,3])"""
    if box is None:
        return "none", None
    if isinstance(box, tuple):      # already classified (tests/fake_backend.py)
        return box
    box = np.asarray(box, dtype=np.float32, order="C")
    if box.shape != (6,):
        raise ValueError("Invalid box information. Must be of the form "
                         "[lx, ly, lz, alpha, beta, gamma].")
    if np.all(box[3:] == 90.0):
        return "ortho", box[:3].copy()
    return "tri_vecs", triclinic_vectors(box)


def _wrap_triclinic(coords, m):
    """_triclinic_pbc restated (exact upstream op order NOT trusted; identity
    for in-cell coordinates).  coords float32 [n,3], m float32 [3,3]."""
    out = coords.copy()
    b = m.astype(np.float64).ravel()
    bx_by = b[3] / b[4]
    cy_cz = b[7] / b[8]
    a_zf = (b[6] - b[3] * b[7] / b[4]) / b[8]
    for i in range(len(out)):
        c0, c1, c2 = (float(v) for v in out[i])
        moved = False
        if c2 < 0.0 or c2 >= b[8]:
            s = np.floor(c2 / b[8])
            c2, c1, c0 = c2 - s * b[8], c1 - s * b[7], c0 - s * b[6]
            moved = True
        lb = c2 * cy_cz
        if c1 < lb or c1 >= lb + b[4]:
            s = np.floor((c1 - lb) / b[4])
            c1, c0 = c1 - s * b[4], c0 - s * b[3]
            moved = True
        lb = c1 * bx_by + c2 * a_zf
        if c0 < lb or c0 >= lb + b[0]:
            s = np.floor((c0 - lb) / b[0])
            c0 = c0 - s * b[0]
            moved = True
        if moved:
            out[i] = (np.float32(c0), np.float32(c1), np.float32(c2))
    return out


def _round_half_away(x):
    """C round(): halves away from zero (np.round is half-to-even)."""
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def distance_array(ref, conf, box=None):
    """MDAnalysis.lib.distances.distance_array restated -> float64 [n1, n2]."""
    ref = np.ascontiguousarray(ref, dtype=np.float32)
    conf = np.ascontiguousarray(conf, dtype=np.float32)
    kind, b = check_box(box)
    if kind == "tri_vecs":
        ref = _wrap_triclinic(ref, b)
        conf = _wrap_triclinic(conf, b)
    # FLOAT32 subtraction conf[j] - ref[i], then widened to double
    dx = (conf[None, :, :] - ref[:, None, :]).astype(np.float64)
    dx = _minimum_image(dx, kind, b)
    rsq = (dx[..., 0] * dx[..., 0]) + (dx[..., 1] * dx[..., 1]) \
        + (dx[..., 2] * dx[..., 2])
    return np.sqrt(rsq)


def _minimum_image(dx, kind, b):
    """minimum_image / minimum_image_triclinic (calc_distances.h) applied to
    float64 difference vectors dx[..., 3] of (already wrapped, if triclinic)
    float32 coordinates."""
    dx = np.array(dx, dtype=np.float64)
    if kind == "ortho":
        inv = (1.0 / b.astype(np.float64)).astype(np.float32)
        for k in range(3):
            if b[k] > FLT_EPSILON:
                s = np.float64(inv[k]) * dx[..., k]
                dx[..., k] = np.float64(b[k]) * (s - _round_half_away(s))
    elif kind == "tri_vecs":
        bf = b.ravel()  # float32; box[k]*i is a float product (exact for +-1)
        best = np.full(dx.shape[:-1], FLT_MAX, dtype=np.float64)
        best_v = np.zeros_like(dx)
        for ix in (-1, 0, 1):
            rx = dx[..., 0] + np.float64(bf[0] * np.float32(ix))
            for iy in (-1, 0, 1):
                ry0 = rx + np.float64(bf[3] * np.float32(iy))
                ry1 = dx[..., 1] + np.float64(bf[4] * np.float32(iy))
                for iz in (-1, 0, 1):
                    rz0 = ry0 + np.float64(bf[6] * np.float32(iz))
                    rz1 = ry1 + np.float64(bf[7] * np.float32(iz))
                    rz2 = dx[..., 2] + np.float64(bf[8] * np.float32(iz))
                    dsq = rz0 * rz0 + rz1 * rz1 + rz2 * rz2
                    upd = dsq < best
                    best = np.where(upd, dsq, best)
                    best_v[..., 0] = np.where(upd, rz0, best_v[..., 0])
                    best_v[..., 1] = np.where(upd, rz1, best_v[..., 1])
                    best_v[..., 2] = np.where(upd, rz2, best_v[..., 2])
        dx = best_v
    return dx


def capped_distance(ref, conf, max_cutoff, box=None):
    """_bruteforce_capped restated -> (pairs intp [P,2], dist float64 [P])."""
    d = distance_array(ref, conf, box=box)
    mask = np.where(d <= max_cutoff)
    pairs = np.c_[mask[0], mask[1]].astype(np.intp)
    return pairs, d[mask]


def rdf_single_frame(g1, g2, dimensions, nbins, rng, exclusion_block=None):
    """InterRDF._single_frame (pmda/rdf.py:120-137) -> (count int64, volume)."""
    pairs, dist = capped_distance(g1, g2, rng[1], box=dimensions)
    if exclusion_block is not None:
        idx_a = pairs[:, 0] // exclusion_block[0]
        idx_b = pairs[:, 1] // exclusion_block[1]
        dist = dist[np.where(idx_a != idx_b)[0]]
    count = np.histogram(dist, bins=nbins, range=rng)[0]
    return count, box_volume(dimensions)


def rdf_s_single_frame(site_pairs, dimensions, nbins, rng):
    """InterRDF_s._single_frame (pmda/rdf.py:292-309).
    site_pairs: list of (pos1 [n1,3], pos2 [n2,3]) -> list of float64 tensors."""
    out = []
    for p1, p2 in site_pairs:
        cnt = np.zeros((len(p1), len(p2), nbins), dtype=np.float64)
        pairs, dist = capped_distance(p1, p2, rng[1], box=dimensions)
        for j, (i1, i2) in enumerate(pairs):
            cnt[i1, i2, :] = np.histogram(dist[j], bins=nbins, range=rng)[0]
        out.append(cnt)
    return out, box_volume(dimensions)


# ---------------------------------------------------------------------------
# Contacts (pmda/contacts.py:270-289).  TEST INFRASTRUCTURE like the rest of
# this module.
# ---------------------------------------------------------------------------
def distance_array_nobox(ref, conf):
    """MDAnalysis _calc_distance_array: float32 difference conf[j] - ref[i],
    double squares summed x, y, z, double sqrt."""
    ref = np.ascontiguousarray(ref, dtype=np.float32)
    conf = np.ascontiguousarray(conf, dtype=np.float32)
    out = np.empty((len(ref), len(conf)), dtype=np.float64)
    for i in range(len(ref)):
        dx = (conf - ref[i]).astype(np.float64)
        out[i] = np.sqrt((dx[:, 0] * dx[:, 0] + dx[:, 1] * dx[:, 1]) +
                         dx[:, 2] * dx[:, 2])
    return out


def contacts_single_frame(frame_no, posA, posB, initial_contacts, r0s,
                          fraction_contacts, fraction_kwargs):
    """Contacts._single_frame restated (pmda/contacts.py:270-289)"""
    d = distance_array_nobox(posA, posB)
    y = np.empty(len(r0s) + 1)
    y[0] = frame_no
    for i, (init, r0) in enumerate(zip(initial_contacts, r0s)):
        y[i + 1] = fraction_contacts(d[init], r0[init], **fraction_kwargs)
    if len(y) == 1:
        y = y[0]
    return y


# ---------------------------------------------------------------------------
# HydrogenBondAnalysis (pmda/hbond_analysis.py:415-470).  TEST INFRASTRUCTURE.
# ---------------------------------------------------------------------------
def capped_distance_minmax(ref, conf, max_cutoff, min_cutoff=None, box=None,
                           distance_fn=None):
    """capped_distance(..., min_cutoff=...) through _bruteforce_capped
    [recalled-upstream lib/distances.py]: ``d <= max_cutoff`` then
    ``d > min_cutoff``; pairs in the row-major order of np.where."""
    fn = distance_array if distance_fn is None else distance_fn
    d = fn(ref, conf, box=box)
    keep = d <= max_cutoff
    if min_cutoff is not None:
        keep &= d > min_cutoff
    mask = np.where(keep)
    pairs = np.c_[mask[0], mask[1]].astype(np.intp)
    return pairs, d[mask]


def calc_angles(a, b, c, box=None):
    """MDAnalysis.lib.distances.calc_angles -> radians float64 [n]: angle at
    the apex b between b->a and b->c.  C loops of calc_distances.h
    (_calc_angle / _calc_angle_ortho / _calc_angle_triclinic)
    [recalled-upstream]: rji = a - b and rjk = c - b are FLOAT32 differences
    widened to double, minimum-imaged like a distance vector (triclinic: of the
    wrapped coordinates), then atan2(|rji x rjk|, rji . rjk)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    c = np.ascontiguousarray(c, dtype=np.float32)
    kind, bx = check_box(box)
    if kind == "tri_vecs":
        a, b, c = (_wrap_triclinic(v, bx) for v in (a, b, c))
    rji = _minimum_image((a - b).astype(np.float64), kind, bx)
    rjk = _minimum_image((c - b).astype(np.float64), kind, bx)
    x = (rji[:, 0] * rjk[:, 0] + rji[:, 1] * rjk[:, 1]) + rji[:, 2] * rjk[:, 2]
    xp0 = rji[:, 1] * rjk[:, 2] - rji[:, 2] * rjk[:, 1]
    xp1 = -rji[:, 0] * rjk[:, 2] + rji[:, 2] * rjk[:, 0]
    xp2 = rji[:, 0] * rjk[:, 1] - rji[:, 1] * rjk[:, 0]
    y = np.sqrt((xp0 * xp0 + xp1 * xp1) + xp2 * xp2)
    return np.arctan2(y, x)


def hbond_rows(positions, donors, hydrogens, acceptors, dimensions,
               max_cutoff, min_cutoff, d_h_a_angle, distance_fn=None):
    """the search of HydrogenBondAnalysis._single_frame
    (pmda/hbond_analysis.py:427-452) -> (k, j, dist, angle_deg): row k of the
    paired donor / hydrogen lists bonded to acceptor j (positions in the
    lists), in the reference's order.  `hydrogens` None: capped_distance only
    (angle 0)."""
    positions = np.asarray(positions, dtype=np.float32)
    donors = np.asarray(donors, dtype=np.intp)
    acceptors = np.asarray(acceptors, dtype=np.intp)
    pairs, dist = capped_distance_minmax(
        positions[donors], positions[acceptors], max_cutoff, min_cutoff,
        box=dimensions, distance_fn=distance_fn)
    k, j = pairs[:, 0], pairs[:, 1]
    if hydrogens is None:
        return k, j, dist, np.zeros(len(k))
    hydrogens = np.asarray(hydrogens, dtype=np.intp)
    ang = np.rad2deg(calc_angles(positions[donors[k]], positions[hydrogens[k]],
                                 positions[acceptors[j]], box=dimensions))
    keep = np.where(ang > d_h_a_angle)[0]
    return k[keep], j[keep], dist[keep], ang[keep]


def hbonds_single_frame(frame_no, positions, donors, hydrogens, acceptors,
                        dimensions, d_a_cutoff=3.0, d_h_a_angle=150.0,
                        min_cutoff=1.0, distance_fn=None):
    """HydrogenBondAnalysis._single_frame restated
    (pmda/hbond_analysis.py:415-470) -> float64 [n_hbonds, 6]:
    frame, donor index, hydrogen index, acceptor index, distance, angle.
    `donors` / `hydrogens` are the paired index lists of _get_dh_pairs."""
    donors, hydrogens, acceptors = (np.asarray(v, dtype=np.intp)
                                    for v in (donors, hydrogens, acceptors))
    k, j, dist, ang = hbond_rows(positions, donors, hydrogens, acceptors,
                                 dimensions, d_a_cutoff, min_cutoff,
                                 d_h_a_angle, distance_fn)
    out = np.empty((len(k), 6), dtype=np.float64)
    out[:, 0] = frame_no
    out[:, 1] = donors[k]
    out[:, 2] = hydrogens[k]
    out[:, 3] = acceptors[j]
    out[:, 4] = dist
    out[:, 5] = ang
    return out
