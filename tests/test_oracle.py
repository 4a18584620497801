"""Pin the oracle where it can be pinned:

* binning      -- oracle_np_bin against the real numpy.histogram the reference
                  calls (pmda/rdf.py:134, :304-305), on values sitting on and
                  within a few ulp of every edge;
* C vs numpy   -- the C restatement against the literal numpy restatement
                  (distance_array -> mask -> np.histogram) bit for bit;
* box helpers  -- product mdamath vs the oracle's independent restatement.
The distance arithmetic itself is restated from MDAnalysis (not available
offline): parity UNPINNED, anchored by tests/test_kat.py.
"""
import numpy as np
import pytest

from oracle import c_oracle, numpy_oracle
from pmda_b200 import mdamath


@pytest.mark.parametrize("nbins,rng", [(75, (0.0, 15.0)), (200, (0.0, 15.0)),
                                       (412, (1.0, 13.0)), (7, (0.3, 0.9)),
                                       (1, (0.0, 2.5))])
def test_np_bin_matches_numpy_histogram(nbins, rng):
    edges = np.histogram([-1], bins=nbins, range=rng)[1]
    vals = [edges]
    for k in range(1, 4):
        up, dn = edges.copy(), edges.copy()
        for _ in range(k):
            up = np.nextafter(up, np.inf)
            dn = np.nextafter(dn, -np.inf)
        vals += [up, dn]
    gen = np.random.default_rng(5)
    vals.append(gen.uniform(rng[0] - 1.0, rng[1] + 1.0, 4000))
    vals.append(np.array([0.0, -0.0, rng[0], rng[1], np.inf, np.nan]))
    vals = np.concatenate(vals)
    for v in vals:
        want = np.histogram(v, bins=nbins, range=rng)[0]
        got = c_oracle.np_bin(v, nbins, rng, edges)
        if want.sum() == 0:
            assert got == -1, v
        else:
            assert got == int(np.argmax(want)), v


BOXES = [
    [20.0, 22.0, 24.0, 90.0, 90.0, 90.0],
    [20.0, 22.0, 24.0, 60.0, 70.0, 80.0],
    [30.0, 30.0, 30.0, 60.0, 60.0, 90.0],
    [25.0, 25.0, 0.0, 90.0, 90.0, 90.0],
    None,
]


def _coords(dims, n, gen, spread=1.0):
    if dims is None or dims[3] == dims[4] == dims[5] == 90.0:
        L = np.array([20.0, 20.0, 20.0] if dims is None else
                     [d if d > 0 else 20.0 for d in dims[:3]])
        return ((gen.random((n, 3)) * spread - (spread - 1) / 2) * L).astype(
            np.float32)
    m = numpy_oracle.triclinic_vectors(dims, np.float64)
    frac = 0.001 + 0.998 * gen.random((n, 3))
    return (frac @ m).astype(np.float32)


@pytest.mark.parametrize("dims", BOXES)
def test_c_oracle_equals_numpy_oracle(dims):
    gen = np.random.default_rng(1)
    a, b = _coords(dims, 61, gen), _coords(dims, 87, gen)
    d_c = c_oracle.distance_array(a, b, dims)
    d_n = numpy_oracle.distance_array(a, b, dims)
    assert np.array_equal(d_c, d_n)
    if dims is not None:
        for excl in (None, (3, 4)):
            c_c, v_c = c_oracle.rdf_single_frame(a, b, dims, 75, (0.0, 15.0),
                                                 excl)
            c_n, v_n = numpy_oracle.rdf_single_frame(a, b, dims, 75,
                                                     (0.0, 15.0), excl)
            assert np.array_equal(c_c, c_n)
            assert v_c == v_n


def test_c_oracle_unwrapped_coordinates():
    gen = np.random.default_rng(2)
    for dims in BOXES[:3]:
        a = (gen.random((40, 3)) * 100 - 50).astype(np.float32)
        b = (gen.random((50, 3)) * 100 - 50).astype(np.float32)
        assert np.array_equal(c_oracle.distance_array(a, b, dims),
                              numpy_oracle.distance_array(a, b, dims))


def test_c_oracle_threads_do_not_change_counts():
    gen = np.random.default_rng(3)
    dims = BOXES[0]
    a = _coords(dims, 700, gen)
    c1, _ = c_oracle.rdf_single_frame(a, a, dims, 75, (0.0, 15.0), nthreads=1)
    c4, _ = c_oracle.rdf_single_frame(a, a, dims, 75, (0.0, 15.0), nthreads=4)
    assert np.array_equal(c1, c4)


def test_rdf_s_oracles_agree():
    gen = np.random.default_rng(4)
    dims = BOXES[1]
    p = _coords(dims, 120, gen)
    pairs = [(p[:1], p[10:40]), (p[50:52], p[60:77])]
    c_c, v_c = c_oracle.rdf_s_single_frame(pairs, dims, 75, (0.0, 15.0))
    c_n, v_n = numpy_oracle.rdf_s_single_frame(pairs, dims, 75, (0.0, 15.0))
    for x, y in zip(c_c, c_n):
        assert np.array_equal(x, y)
    assert v_c == v_n


@pytest.mark.parametrize("dims", BOXES[:4] + [[10.0, 11.0, 12.0, 90.0, 60.0,
                                              120.0]])
def test_box_helpers_product_vs_oracle(dims):
    assert np.array_equal(mdamath.triclinic_vectors(dims),
                          numpy_oracle.triclinic_vectors(dims))
    assert mdamath.box_volume(dims) == numpy_oracle.box_volume(dims)
    kind_o, b_o = numpy_oracle.check_box(dims)
    kind_p, b_p = mdamath.classify_box(dims)
    assert {"ortho": 1, "tri_vecs": 2}[kind_o] == kind_p
    assert np.array_equal(np.asarray(b_o, dtype=np.float32).ravel(),
                          b_p[:np.asarray(b_o).size])


def test_invalid_box_vectors_are_zero():
    assert not mdamath.triclinic_vectors([10, 10, 10, 10, 20, 170]).any()
    assert not mdamath.triclinic_vectors([10, -1, 10, 90, 90, 90]).any()
    assert mdamath.box_volume([10, 10, 10, 10, 20, 170]) == 0.0


@pytest.mark.parametrize("L,n,rmax", [
    ((40.0, 44.0, 38.0), 3000, 9.0),     # several cells per axis
    ((20.0, 50.0, 30.0), 2500, 9.5),     # two / five / three cells
    ((16.0, 17.0, 15.5), 1200, 9.0),     # one cell per axis: every pair
])
def test_cell_list_oracle_equals_full_pair_matrix(L, n, rmax):
    """oracle_rdf_frame_cells (the faster CPU baseline) must give the integers
    of the O(N1 N2) oracle, also for unwrapped coordinates and exclusions"""
    gen = np.random.default_rng(31)
    Lv = np.array(L)
    pos = (gen.random((n, 3)) * Lv * 3 - Lv).astype(np.float32)   # unwrapped
    dims = list(L) + [90, 90, 90]
    a, b = pos[: n // 3], pos
    for excl in (None, (3, 5)):
        full, _ = c_oracle.rdf_single_frame(a, b, dims, 60, (0.0, rmax), excl,
                                            nthreads=4)
        cell, _ = c_oracle.rdf_single_frame(a, b, dims, 60, (0.0, rmax), excl,
                                            nthreads=4, cells=True)
        np.testing.assert_array_equal(cell, full)
    # partial PBC: no cell list, silently the full matrix
    dims0 = [L[0], L[1], 0.0, 90, 90, 90]
    full, _ = c_oracle.rdf_single_frame(a, b, dims0, 40, (0.0, rmax))
    cell, _ = c_oracle.rdf_single_frame(a, b, dims0, 40, (0.0, rmax), cells=True)
    np.testing.assert_array_equal(cell, full)
