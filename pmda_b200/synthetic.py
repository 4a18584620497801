"""Synthetic trajectories for the BASELINE.json configurations.

Everything is drawn from ``numpy.random.Generator(PCG64(seed))`` in float64,
cast to float32 and kept strictly inside the primary cell (1e-4 fractional
margin for triclinic boxes) so that every plausible pre-wrap of the reference
is the identity (SURVEY.md 8c "residual risk").
"""
import numpy as np

from . import mdamath
from .universe import Universe

BASE_SEED = 20261019


def _rng(cfg, extra=0):
    return np.random.Generator(np.random.PCG64(BASE_SEED + cfg + 1000 * extra))


def _clamp_ortho(pos, lengths):
    hi = np.nextafter(np.asarray(lengths, dtype=np.float32), np.float32(0))
    return np.minimum(np.maximum(pos, np.float32(0)), hi)


def uniform_ortho(n_frames, n_atoms, lengths, rng):
    """ideal gas in an orthorhombic box -> float32 [F,N,3]"""
    L = np.asarray(lengths, dtype=np.float64)
    pos = (rng.random((n_frames, n_atoms, 3)) * L).astype(np.float32)
    return _clamp_ortho(pos, L)


def uniform_triclinic(n_frames, n_atoms, dimensions, rng, margin=1e-4):
    """ideal gas inside a triclinic cell (fractional coords in
    [margin, 1-margin]) -> float32 [F,N,3]"""
    m = mdamath.triclinic_vectors(dimensions, dtype=np.float64)
    frac = margin + (1.0 - 2.0 * margin) * rng.random((n_frames, n_atoms, 3))
    return (frac @ m).astype(np.float32)


def water_like(n_frames, n_atoms, L, rng, sigma=0.5):
    """jittered simple-cubic lattice + per-frame Gaussian displacement"""
    m = int(np.ceil(n_atoms ** (1.0 / 3.0)))
    a = L / m
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"),
                 axis=-1).reshape(-1, 3)
    sites = (g[rng.permutation(len(g))[:n_atoms]] + 0.5) * a
    pos = sites[None] + rng.normal(0.0, sigma, (n_frames, n_atoms, 3))
    pos = np.mod(pos, L)
    return _clamp_ortho(pos.astype(np.float32), [L, L, L])


def water_box(n_frames, n_mol, dimensions, rng, jitter=0.05):
    """rigid three-site waters (O, H1, H2; O-H 0.9572 A, H-O-H 104.52 deg) at
    random positions and orientations in the cell -> (float32 positions
    [n_frames, 3 n_mol, 3], names).  Atom order O H H per molecule.  Molecules
    are not wrapped as a whole: hydrogens may lie in a neighbouring image of
    their oxygen, which is what the minimum-image angle code has to handle."""
    from .mdamath import triclinic_vectors
    dims = np.asarray(dimensions, dtype=np.float64)
    m = np.asarray(triclinic_vectors(dims), dtype=np.float64)
    frac = rng.random((n_mol, 3))
    centre = frac @ m
    out = np.empty((n_frames, n_mol, 3, 3), dtype=np.float64)
    half = np.deg2rad(104.52) / 2.0
    local = 0.9572 * np.array([[np.sin(half), np.cos(half), 0.0],
                               [-np.sin(half), np.cos(half), 0.0]])
    for f in range(n_frames):
        q = rng.normal(size=(n_mol, 4))
        q /= np.linalg.norm(q, axis=1)[:, None]
        w, x, y, z = q.T
        R = np.stack([
            np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
            np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
            np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1),
        ], 1)
        o = centre + rng.normal(0.0, jitter * (f + 1), (n_mol, 3))
        # wrap the oxygens into the cell through fractional coordinates
        fo = np.mod(o @ np.linalg.inv(m), 1.0)
        fo = np.clip(fo, 1e-5, 1.0 - 1e-5)
        o = fo @ m
        out[f, :, 0] = o
        out[f, :, 1] = o + np.einsum("nij,j->ni", R, local[0])
        out[f, :, 2] = o + np.einsum("nij,j->ni", R, local[1])
    names = np.tile(np.array(["OW", "HW1", "HW2"]), n_mol)
    return out.reshape(n_frames, 3 * n_mol, 3).astype(np.float32), names


# -- the five BASELINE.json configurations ----------------------------------
def config1(n_frames=100, n_atoms=3000, seed_extra=0):
    """InterRDF O-O, 3,000-water box, cubic L=44.78, 75 bins, range 0-15"""
    L = 44.78
    rng = _rng(1, seed_extra)
    pos = water_like(n_frames, n_atoms, L, rng)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    return u, dict(g1=u.atoms, g2=u.atoms, nbins=75, range=(0.0, 15.0))


def config2(n_frames=1000, n_atoms=100_000, seed_extra=0):
    """InterRDF all-atom, 100k atoms, cubic L=100, 200 bins, range 0-15"""
    L = 100.0
    rng = _rng(2, seed_extra)
    pos = uniform_ortho(n_frames, n_atoms, [L, L, L], rng)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    return u, dict(g1=u.atoms, g2=u.atoms, nbins=200, range=(0.0, 15.0))


def config3(n_frames=2000, n_atoms=50_000, n_sites=64, n_near=256,
            seed_extra=0):
    """InterRDF_s: 64 ion-water site pairs (1 ion x 256 nearest atoms of
    frame 0) in a 50k-atom cubic box L=79.4, 75 bins, range 0-15"""
    L = 79.4
    rng = _rng(3, seed_extra)
    base = uniform_ortho(1, n_atoms, [L, L, L], rng)[0].astype(np.float64)
    pos = base[None] + rng.normal(0.0, 0.3, (n_frames, n_atoms, 3))
    pos = _clamp_ortho(np.mod(pos, L).astype(np.float32), [L, L, L])
    u = Universe(pos, [L, L, L, 90, 90, 90])
    ions = rng.choice(n_atoms, n_sites, replace=False)
    ags = []
    p0 = pos[0].astype(np.float64)
    for ion in ions:
        d = p0 - p0[ion]
        d -= L * np.round(d / L)
        order = np.argsort(np.einsum("ij,ij->i", d, d))
        near = np.sort(order[1:n_near + 1])
        ags.append([u.atoms[[ion]], u.atoms[near]])
    return u, dict(ags=ags, nbins=75, range=(0.0, 15.0), density=True)


def config4(n_frames=5000, n_atoms=250_000, seed_extra=0):
    """InterRDF, 250k atoms, rhombic-dodecahedron-like triclinic box"""
    dims = [152.3, 152.3, 152.3, 60.0, 60.0, 90.0]
    rng = _rng(4, seed_extra)
    pos = uniform_triclinic(n_frames, n_atoms, dims, rng)
    u = Universe(pos, dims)
    return u, dict(g1=u.atoms, g2=u.atoms, nbins=75, range=(0.0, 15.0))


def config5(n_frames=10_000, n_atoms=1_000_000, n_g1=1000, seed_extra=0):
    """InterRDF 1k x 1M, membrane-like slab density, ortho 316.2x316.2x100"""
    Ls = [316.2, 316.2, 100.0]
    rng = _rng(5, seed_extra)
    pos = uniform_ortho(n_frames, n_atoms, Ls, rng)
    # half of the atoms sit in two Gaussian slabs (z = 35 / 65, sigma = 8)
    n_slab = n_atoms // 2
    centre = np.where(rng.random(n_slab) < 0.5, 35.0, 65.0)
    z = centre[None] + rng.normal(0.0, 8.0, (n_frames, n_slab))
    pos[:, :n_slab, 2] = np.mod(z, Ls[2]).astype(np.float32)
    pos = _clamp_ortho(pos, Ls)
    u = Universe(pos, Ls + [90, 90, 90])
    g1 = u.atoms[np.sort(rng.choice(n_slab, n_g1, replace=False))]
    return u, dict(g1=g1, g2=u.atoms, nbins=75, range=(0.0, 15.0))


CONFIGS = {1: config1, 2: config2, 3: config3, 4: config4, 5: config5}


# -- the same five configurations, generated frame by frame ------------------
# Every frame has its own seed, so each rank of a multi-GPU run can generate
# just the frames of its shard (bench.py): `only` = the frame numbers to fill;
# the other frames stay uninitialised memory that the rank never reads.
def _frame_rng(cfg, f):
    return np.random.Generator(np.random.PCG64([BASE_SEED, cfg, 1 + int(f)]))


def _static_rng(cfg):
    return np.random.Generator(np.random.PCG64([BASE_SEED, cfg, 0]))


def sharded_config(cfg, n_frames, only=None, n_atoms=None):
    """-> (Universe, kwargs) like ``CONFIGS[cfg]`` but frame-seeded.  The
    static part (lattice sites, site selections, slab membership) comes from a
    per-config seed and is identical on every rank."""
    frames = range(n_frames) if only is None else only
    srng = _static_rng(cfg)
    if cfg == 1:
        n, L = n_atoms or 3000, 44.78
        m = int(np.ceil(n ** (1.0 / 3.0)))
        g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"),
                     axis=-1).reshape(-1, 3)
        sites = (g[srng.permutation(len(g))[:n]] + 0.5) * (L / m)
        pos = np.empty((n_frames, n, 3), dtype=np.float32)
        for f in frames:
            p = np.mod(sites + _frame_rng(cfg, f).normal(0.0, 0.5, (n, 3)), L)
            pos[f] = _clamp_ortho(p.astype(np.float32), [L, L, L])
        u = Universe(pos, [L, L, L, 90, 90, 90])
        return u, dict(g1=u.atoms, g2=u.atoms, nbins=75, range=(0.0, 15.0))
    if cfg == 2:
        n, L = n_atoms or 100_000, 100.0
        pos = np.empty((n_frames, n, 3), dtype=np.float32)
        for f in frames:
            pos[f] = uniform_ortho(1, n, [L, L, L], _frame_rng(cfg, f))[0]
        u = Universe(pos, [L, L, L, 90, 90, 90])
        return u, dict(g1=u.atoms, g2=u.atoms, nbins=200, range=(0.0, 15.0))
    if cfg == 3:
        n, L, n_sites, n_near = n_atoms or 50_000, 79.4, 64, 256
        base = uniform_ortho(1, n, [L, L, L], srng)[0].astype(np.float64)
        pos = np.empty((n_frames, n, 3), dtype=np.float32)
        for f in frames:
            p = np.mod(base + _frame_rng(cfg, f).normal(0.0, 0.3, (n, 3)), L)
            pos[f] = _clamp_ortho(p.astype(np.float32), [L, L, L])
        u = Universe(pos, [L, L, L, 90, 90, 90])
        ions = srng.choice(n, n_sites, replace=False)
        ags = []
        for ion in ions:          # the 256 atoms nearest to the ion's site
            d = base - base[ion]
            d -= L * np.round(d / L)
            order = np.argsort(np.einsum("ij,ij->i", d, d))
            ags.append([u.atoms[[ion]], u.atoms[np.sort(order[1:n_near + 1])]])
        return u, dict(ags=ags, nbins=75, range=(0.0, 15.0), density=True)
    if cfg == 4:
        n, dims = n_atoms or 250_000, [152.3, 152.3, 152.3, 60.0, 60.0, 90.0]
        pos = np.empty((n_frames, n, 3), dtype=np.float32)
        for f in frames:
            pos[f] = uniform_triclinic(1, n, dims, _frame_rng(cfg, f))[0]
        u = Universe(pos, dims)
        return u, dict(g1=u.atoms, g2=u.atoms, nbins=75, range=(0.0, 15.0))
    if cfg == 5:
        n, Ls, n_g1 = n_atoms or 1_000_000, [316.2, 316.2, 100.0], 1000
        n_slab = n // 2
        centre = np.where(srng.random(n_slab) < 0.5, 35.0, 65.0)
        pos = np.empty((n_frames, n, 3), dtype=np.float32)
        for f in frames:
            rng = _frame_rng(cfg, f)
            p = uniform_ortho(1, n, Ls, rng)[0]
            z = centre + rng.normal(0.0, 8.0, n_slab)
            p[:n_slab, 2] = np.mod(z, Ls[2]).astype(np.float32)
            pos[f] = _clamp_ortho(p, Ls)
        u = Universe(pos, Ls + [90, 90, 90])
        g1 = u.atoms[np.sort(srng.choice(n_slab, min(n_g1, n_slab),
                                         replace=False))]
        return u, dict(g1=g1, g2=u.atoms, nbins=75, range=(0.0, 15.0))
    raise ValueError("no BASELINE config %r" % (cfg,))
