"""Known-answer tests for the oracle: the distance arithmetic is
restated from MDAnalysis and cannot be pinned against it offline, so these fix
its observable behaviour from first principles."""
import numpy as np
import pytest

from oracle import c_oracle, numpy_oracle


def lattice(m, a):
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"),
                 axis=-1).reshape(-1, 3)
    return (g * a).astype(np.float32)


# neighbours of a simple cubic lattice by squared distance / a^2
SC_SHELLS = {1: 6, 2: 12, 3: 8, 4: 6, 5: 24, 6: 24, 8: 12, 9: 30}


def test_simple_cubic_shell_counts():
    """every atom sees exactly 6, 12, 8, 6, 24, 24, 12, 30 neighbours at
    a*sqrt(1, 2, 3, 4, 5, 6, 8, 9) -- distances kept away from bin edges"""
    m, a = 8, 1.3037
    pos = lattice(m, a)
    L = m * a
    dims = [L, L, L, 90, 90, 90]
    nbins, rng = 400, (0.0, 4.0)
    count, _ = c_oracle.rdf_single_frame(pos, pos, dims, nbins, rng)
    edges = np.histogram([-1], bins=nbins, range=rng)[1]
    n = len(pos)
    assert count[0] == n                      # self pairs, d == 0 (quirk 1)
    for s2, mult in SC_SHELLS.items():
        d = a * np.sqrt(s2)
        k = np.searchsorted(edges, d, side="right") - 1
        assert min(d - edges[k], edges[k + 1] - d) > 1e-4
        assert count[k] == n * mult, (s2, count[k])
    assert count.sum() == n * (1 + sum(SC_SHELLS.values()))


def test_lattice_on_edges_follows_numpy_rule():
    """spacing = integer number of bin widths: shells at a, 2a, 3a sit exactly
    on edges (up to the float32-inverse-box perturbation); the C oracle must
    agree with distance_array + np.histogram bit for bit"""
    m, a = 8, 1.0
    pos = lattice(m, a)
    L = m * a
    dims = [L, L, L, 90, 90, 90]
    for nbins, rng in ((20, (0.0, 4.0)), (15, (0.0, 3.0)), (40, (0.0, 4.0))):
        c_c, _ = c_oracle.rdf_single_frame(pos, pos, dims, nbins, rng)
        c_n, _ = numpy_oracle.rdf_single_frame(pos, pos, dims, nbins, rng)
        assert np.array_equal(c_c, c_n)
        # the right edge of the range is inclusive twice over (quirk 4)
        d = numpy_oracle.distance_array(pos, pos, dims)
        assert c_c.sum() == np.count_nonzero(d <= rng[1])


def test_ideal_gas_rdf_is_one():
    gen = np.random.default_rng(8)
    n, L = 4000, 40.0
    pos = (gen.random((n, 3)) * L).astype(np.float32)
    dims = [L, L, L, 90, 90, 90]
    nbins, rng = 30, (0.0, 15.0)
    count, vol = c_oracle.rdf_single_frame(pos, pos, dims, nbins, rng,
                                           nthreads=8)
    edges = np.histogram([-1], bins=nbins, range=rng)[1]
    shell = 4 / 3 * np.pi * (edges[1:] ** 3 - edges[:-1] ** 3)
    rdf = count / (n * n / vol * shell)
    assert vol == L ** 3
    # skip bin 0 (self pairs); Poisson error ~ 1/sqrt(count)
    for k in range(3, nbins):
        assert abs(rdf[k] - 1.0) < 6.0 / np.sqrt(count[k]), k
    total = n * n * (4 / 3 * np.pi * 15.0 ** 3) / L ** 3
    assert abs(count.sum() - n - total) < 6 * np.sqrt(total)


@pytest.mark.parametrize("dims", [[30.0, 30.0, 30.0, 90, 90, 90],
                                  [30.0, 31.0, 32.0, 75.0, 80.0, 85.0]])
def test_two_atoms_across_an_edge(dims):
    """sweep one atom across a bin edge in ~1e-7 steps, wrapped through the
    periodic boundary: C oracle == numpy oracle, and the count moves from bin k
    to bin k+1 exactly once"""
    nbins, rng = 75, (0.0, 15.0)
    edges = np.histogram([-1], bins=nbins, range=rng)[1]
    e = edges[20]                           # 4.0
    m = numpy_oracle.triclinic_vectors(dims, np.float64)
    origin = (np.array([0.05, 0.5, 0.5]) @ m)
    seen = []
    for i in range(-40, 41):
        x = e + i * 2.5e-7
        # second atom sits x away through the periodic image along a
        other = origin - np.array([x, 0.0, 0.0]) + m[0]
        a = origin.astype(np.float32)[None]
        b = other.astype(np.float32)[None]
        c_c, _ = c_oracle.rdf_single_frame(a, b, dims, nbins, rng)
        c_n, _ = numpy_oracle.rdf_single_frame(a, b, dims, nbins, rng)
        assert np.array_equal(c_c, c_n)
        assert c_c.sum() == 1
        seen.append(int(np.argmax(c_c)))
    assert seen[0] == 19 and seen[-1] == 20
    assert all(b >= a for a, b in zip(seen, seen[1:]))


def test_exclusion_block_removes_whole_tiles():
    gen = np.random.default_rng(9)
    n, k, L = 600, 3, 9.0                 # box small enough: all pairs in range
    pos = (gen.random((n, 3)) * L).astype(np.float32)
    dims = [L, L, L, 90, 90, 90]
    c0, _ = c_oracle.rdf_single_frame(pos, pos, dims, 75, (0.0, 15.0))
    c1, _ = c_oracle.rdf_single_frame(pos, pos, dims, 75, (0.0, 15.0), (k, k))
    assert c0.sum() == n * n
    assert c0.sum() - c1.sum() == n * k     # N * k in-range self-block pairs


def test_ortho_matches_textbook_minimum_image_loosely():
    """sanity: within 1e-4 A of |dx - L*round(dx/L)| (not bit-equal: the
    reference uses a float32-rounded inverse box)"""
    gen = np.random.default_rng(10)
    L = np.array([31.0, 17.0, 23.0])
    a = (gen.random((50, 3)) * L).astype(np.float32)
    b = (gen.random((60, 3)) * L).astype(np.float32)
    d = c_oracle.distance_array(a, b, list(L) + [90, 90, 90])
    dx = b[None].astype(np.float64) - a[:, None].astype(np.float64)
    dx -= L * np.round(dx / L)
    ref = np.sqrt((dx ** 2).sum(-1))
    assert np.max(np.abs(d - ref)) < 1e-4


def test_triclinic_right_angles_given_as_vectors():
    """a 90/90/90 box is classified ortho (check_box), a 90/90/89.999 one is
    triclinic; both must give nearly the same distances for in-cell atoms"""
    gen = np.random.default_rng(11)
    L = 25.0
    a = (gen.random((40, 3)) * L * 0.98 + 0.2).astype(np.float32)
    d_o = c_oracle.distance_array(a, a, [L, L, L, 90, 90, 90])
    d_t = c_oracle.distance_array(a, a, [L, L, L, 90, 90, 89.9999])
    assert np.max(np.abs(d_o - d_t)) < 1e-3


def _hist_agrees_up_to_edge_ties(count, dist64, edges, tol):
    """`count` (oracle, float32-flavoured arithmetic) against the histogram of
    independently computed float64 distances: they may differ only by pairs
    that sit within `tol` of a bin edge."""
    ind = np.histogram(dist64, bins=edges)[0]
    near = np.array([np.sum(np.abs(dist64 - e) < tol) for e in edges])
    slack = near[:-1] + near[1:]
    assert np.all(np.abs(count - ind) <= slack), (count - ind, slack)
    assert abs(int(count.sum()) - int(ind.sum())) <= near[0] + near[-1]


def test_ortho_histogram_against_scipy_periodic_kdtree():
    """independent implementation: scipy's periodic cKDTree (float64,
    true minimum image) must give the oracle's histogram up to edge ties"""
    from scipy.spatial import cKDTree
    gen = np.random.default_rng(12)
    L = np.array([28.0, 31.0, 26.0])
    a = (gen.random((700, 3)) * L).astype(np.float32)
    b = (gen.random((900, 3)) * L).astype(np.float32)
    hi = np.nextafter(L.astype(np.float32), np.float32(0))
    a, b = np.minimum(a, hi), np.minimum(b, hi)
    nbins, rng = 60, (0.0, 12.0)
    count, _ = c_oracle.rdf_single_frame(a, b, list(L) + [90, 90, 90], nbins, rng)
    ta = cKDTree(a.astype(np.float64), boxsize=L)
    tb = cKDTree(b.astype(np.float64), boxsize=L)
    sp = ta.sparse_distance_matrix(tb, rng[1] + 1e-3, output_type="ndarray")
    edges = np.histogram([-1], bins=nbins, range=rng)[1]
    _hist_agrees_up_to_edge_ties(count, sp["v"], edges, 1e-4)


@pytest.mark.parametrize("dims", [
    [30.0, 32.0, 34.0, 80.0, 75.0, 70.0],
    [36.0, 36.0, 36.0, 60.0, 60.0, 90.0],
])
def test_triclinic_histogram_against_lattice_image_search(dims):
    """independent implementation: float64 minimum over 125 lattice images
    (a superset of the reference's 27) for atoms inside the cell"""
    from pmda_b200 import mdamath, synthetic
    gen = np.random.default_rng(13)
    pos = synthetic.uniform_triclinic(1, 500, dims, gen)[0]
    a, b = pos[:200], pos[200:]
    nbins, rng = 50, (0.0, 10.0)
    count, _ = c_oracle.rdf_single_frame(a, b, dims, nbins, rng)
    m = mdamath.triclinic_vectors(dims, dtype=np.float64)
    dx = b[None].astype(np.float64) - a[:, None].astype(np.float64)
    best = np.full(dx.shape[:2], np.inf)
    for i in range(-2, 3):
        for j in range(-2, 3):
            for k in range(-2, 3):
                s = dx + i * m[0] + j * m[1] + k * m[2]
                best = np.minimum(best, np.einsum("ijk,ijk->ij", s, s))
    d = np.sqrt(best).ravel()
    edges = np.histogram([-1], bins=nbins, range=rng)[1]
    _hist_agrees_up_to_edge_ties(count, d[d <= rng[1] + 1e-3], edges, 1e-4)
