"""GPU parity for the sorted / pruned pair kernel: everything the spatial sort,
the box tests and the one-shift-per-group fast path could get wrong must still
give the oracle's integers.  rdfk_set_debug_flags (bit 0: no pruning, bit 1: no shifted
variant, bit 2: no Euclidean refinement of the triclinic box test, bit 4: the
general candidate-bin arithmetic although the range starts at 0) is read by
librdfk on every call;
results must not depend on it."""
import os

import numpy as np
import pytest

from pmda_b200 import _cabi, synthetic
from pmda_b200.rdf import InterRDF
from pmda_b200.universe import Universe

from helpers import oracle_interrdf

pytestmark = pytest.mark.gpu


def _gpu(u, g1, g2, nbins=75, rng=(0.0, 15.0), excl=None, debug=None):
    old = _cabi.set_debug_flags(0 if debug is None else int(debug))
    try:
        r = InterRDF(g1, g2, nbins=nbins, range=rng, exclusion_block=excl)
        r.run()
        return r.count.copy()
    finally:
        _cabi.set_debug_flags(old)


def _check_all_modes(u, g1, g2, nbins=75, rng=(0.0, 15.0), excl=None):
    want, _, _ = oracle_interrdf(u, g1, g2, nbins, rng, excl)
    for debug in (None, 1, 2, 3, 4, 6, 16, 19):
        got = _gpu(u, g1, g2, nbins, rng, excl, debug)
        np.testing.assert_array_equal(got, want, err_msg="debug flags %r" % debug)
    return want


def test_midsize_ortho_all_modes():
    # big enough that most group pairs are pruned and most of the rest shifted
    rng = np.random.default_rng(101)
    L = np.array([62.0, 58.0, 66.0])
    pos = (rng.random((2, 20000, 3)) * L).astype(np.float32)
    u = Universe(pos, list(L) + [90, 90, 90])
    want = _check_all_modes(u, u.atoms, u.atoms, nbins=200)
    assert want[0] >= 2 * 20000


def test_two_groups_swapped_roles_all_modes():
    rng = np.random.default_rng(102)
    L = 55.0
    pos = (rng.random((2, 9000, 3)) * L).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    perm = rng.permutation(9000)
    small, large = u.atoms[np.sort(perm[:500])], u.atoms[np.sort(perm[500:])]
    _check_all_modes(u, small, large)          # i side = g2 (swapped)
    _check_all_modes(u, large, small)          # i side = g1


def test_centered_box_negative_coordinates():
    # coordinates in [-L/2, L/2): the wrap origin is the frame minimum
    rng = np.random.default_rng(103)
    L = 48.0
    pos = ((rng.random((2, 6000, 3)) - 0.5) * L).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    _check_all_modes(u, u.atoms, u.atoms, rng=(0.0, 12.0))


def test_slightly_unwrapped_and_far_unwrapped():
    rng = np.random.default_rng(104)
    L = 40.0
    pos = (rng.random((2, 4000, 3)) * L).astype(np.float32)
    pos[0, :50] += np.float32(L)            # a few atoms one image away
    pos[1, 100:130] -= np.float32(3 * L)    # and three images away
    u = Universe(pos, [L, L, L, 90, 90, 90])
    _check_all_modes(u, u.atoms, u.atoms, rng=(0.0, 10.0))


def test_cutoff_larger_than_half_box():
    rng = np.random.default_rng(105)
    L = np.array([22.0, 26.0, 24.0])
    pos = (rng.random((2, 1500, 3)) * L).astype(np.float32)
    u = Universe(pos, list(L) + [90, 90, 90])
    _check_all_modes(u, u.atoms, u.atoms, rng=(0.0, 15.0))


def test_clustered_density_with_vacuum():
    # blobs + slab: many empty cells, very uneven cluster extents
    rng = np.random.default_rng(106)
    L = np.array([80.0, 80.0, 120.0])
    n = 8000
    centres = rng.random((12, 3)) * L
    blob = centres[rng.integers(0, 12, n // 2)] + rng.normal(0, 3.0, (n // 2, 3))
    slab = rng.random((n - n // 2, 3)) * L
    slab[:, 2] = 60.0 + rng.normal(0, 4.0, n - n // 2)
    p = np.mod(np.concatenate([blob, slab]), L)
    pos = np.stack([p, np.mod(p + rng.normal(0, 0.5, p.shape), L)]).astype(np.float32)
    pos = np.minimum(pos, np.nextafter(L.astype(np.float32), np.float32(0)))
    u = Universe(pos, list(L) + [90, 90, 90])
    _check_all_modes(u, u.atoms, u.atoms, nbins=100)
    _check_all_modes(u, u.atoms[:1000], u.atoms[4000:], nbins=100)


def test_all_atoms_identical_and_tiny_groups():
    L = 30.0
    pos = np.full((1, 300, 3), 7.25, dtype=np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    want = _check_all_modes(u, u.atoms, u.atoms)
    assert want[0] == 300 * 300
    rng = np.random.default_rng(107)
    pos = (rng.random((3, 40, 3)) * L).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    _check_all_modes(u, u.atoms[:1], u.atoms[1:])
    _check_all_modes(u, u.atoms[:7], u.atoms[5:9])


def test_no_box_clustered_all_modes():
    rng = np.random.default_rng(108)
    pos = np.concatenate([rng.normal(0, 6, (1, 3000, 3)),
                          rng.normal(40, 6, (1, 3000, 3))], axis=1)
    u = Universe(pos.astype(np.float32), None)
    _check_all_modes(u, u.atoms, u.atoms)
    _check_all_modes(u, u.atoms[:3000], u.atoms[3000:], rng=(20.0, 90.0))


def test_partial_pbc_all_modes():
    rng = np.random.default_rng(109)
    pos = (rng.random((2, 5000, 3)) * [50, 50, 90]).astype(np.float32)
    u = Universe(pos, [50.0, 50.0, 0.0, 90, 90, 90])
    _check_all_modes(u, u.atoms, u.atoms)


@pytest.mark.parametrize("dims", [
    [60.0, 62.0, 64.0, 80.0, 75.0, 70.0],
    [70.0, 70.0, 70.0, 60.0, 60.0, 90.0],
])
def test_triclinic_midsize_all_modes(dims):
    rng = np.random.default_rng(110)
    pos = synthetic.uniform_triclinic(2, 12000, dims, rng)
    u = Universe(pos, dims)
    _check_all_modes(u, u.atoms, u.atoms, rng=(0.0, 12.0))
    _check_all_modes(u, u.atoms[:300], u.atoms[300:], rng=(0.0, 12.0))


def test_triclinic_exact_frames_all_modes():
    # cut-off beyond the brick: every evaluated pair on the fp64 path
    dims = [24.0, 26.0, 28.0, 70.0, 80.0, 65.0]
    rng = np.random.default_rng(111)
    pos = synthetic.uniform_triclinic(1, 700, dims, rng)
    u = Universe(pos, dims)
    _check_all_modes(u, u.atoms, u.atoms)
    _check_all_modes(u, u.atoms[:100], u.atoms[90:])


def test_nonfinite_coordinates_match_oracle():
    rng = np.random.default_rng(112)
    L = 30.0
    pos = (rng.random((2, 900, 3)) * L).astype(np.float32)
    pos[0, 17, 1] = np.nan
    pos[1, 400, 0] = np.inf
    u = Universe(pos, [L, L, L, 90, 90, 90])
    _check_all_modes(u, u.atoms, u.atoms)
    dims = [40.0, 42.0, 44.0, 80.0, 75.0, 70.0]
    pos = synthetic.uniform_triclinic(1, 600, dims, rng)
    pos[0, 5, 2] = np.nan
    u = Universe(pos, dims)
    _check_all_modes(u, u.atoms, u.atoms, rng=(0.0, 12.0))


def test_large_nbins_global_atomics_and_shared_copies():
    rng = np.random.default_rng(113)
    L = 36.0
    pos = (rng.random((1, 3000, 3)) * L).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    for nbins in (3000, 9000):
        _check_all_modes(u, u.atoms, u.atoms, nbins=nbins, rng=(0.5, 14.0))


def test_box_changes_every_frame():
    rng = np.random.default_rng(114)
    F, n = 5, 5000
    Ls = 50.0 + rng.random((F, 3)) * 4.0
    pos = (rng.random((F, n, 3)) * Ls[:, None, :]).astype(np.float32)
    dims = np.concatenate([Ls, np.full((F, 3), 90.0)], axis=1).astype(np.float32)
    u = Universe(pos, dims)
    _check_all_modes(u, u.atoms, u.atoms)


def test_many_small_batches_with_changing_box():
    """the host enqueues batches far ahead of the device (frame cache hit, two
    compute streams): per-batch device buffers such as the box rows must not be
    recycled while earlier batches still wait to run"""
    from pmda_b200 import engine
    rng = np.random.default_rng(117)
    F, n = 48, 6000
    Ls = 46.0 + rng.random((F, 3)) * 8.0
    pos = (rng.random((F, n, 3)) * Ls[:, None, :]).astype(np.float32)
    dims = np.concatenate([Ls, np.full((F, 3), 90.0)], axis=1).astype(np.float32)
    u = Universe(pos, dims)
    want, _, _ = oracle_interrdf(u, u.atoms, u.atoms)
    old = engine.Config.batch_bytes, engine.Config.cache_bytes
    try:
        engine.Config.batch_bytes = 3 * n * 12          # 3 frames per batch
        for rep in range(3):                            # 2nd, 3rd: cache hits
            got = _gpu(u, u.atoms, u.atoms)
            np.testing.assert_array_equal(got, want, err_msg="rep %d" % rep)
        # a frame cache of two batches: every batch evicts a block whose
        # kernels may still be queued
        engine.clear_cache()
        engine.Config.cache_bytes = 2 * 3 * n * 12
        for rep in range(2):
            got = _gpu(u, u.atoms, u.atoms)
            np.testing.assert_array_equal(got, want, err_msg="evict %d" % rep)
    finally:
        engine.Config.batch_bytes, engine.Config.cache_bytes = old
        engine.clear_cache()


def test_exclusion_block_with_sorted_atoms():
    rng = np.random.default_rng(115)
    L = 45.0
    pos = (rng.random((2, 6000, 3)) * L).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    _check_all_modes(u, u.atoms, u.atoms, excl=(3, 3))
    _check_all_modes(u, u.atoms[:2000], u.atoms[2000:], excl=(2, 4))


def test_lattice_on_edges_all_modes():
    m, a = 16, 1.5
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"),
                 axis=-1).reshape(-1, 3)
    pos = (g * a).astype(np.float32)[None]
    L = m * a
    u = Universe(pos, [L, L, L, 90, 90, 90])
    _check_all_modes(u, u.atoms, u.atoms, nbins=20, rng=(0.0, 6.0))
    _check_all_modes(u, u.atoms, u.atoms, nbins=30, rng=(0.0, 9.0))


def test_large_coordinate_offsets_and_anisotropic_boxes():
    """error bounds scale with the coordinate magnitude: far from the origin
    many more pairs take the fp64 path, the integers stay the oracle's"""
    rng = np.random.default_rng(116)
    L = np.array([40.0, 40.0, 40.0])
    pos = (rng.random((2, 3000, 3)) * L + np.array([5000.0, -7000.0, 12000.0])
           ).astype(np.float32)
    u = Universe(pos, list(L) + [90, 90, 90])
    _check_all_modes(u, u.atoms, u.atoms, rng=(0.0, 12.0))
    # a slab-like cell: one axis barely longer than twice the cut-off
    L = np.array([200.0, 31.0, 25.0])
    pos = (rng.random((2, 4000, 3)) * L).astype(np.float32)
    u = Universe(pos, list(L) + [90, 90, 90])
    _check_all_modes(u, u.atoms, u.atoms, rng=(0.0, 12.0))
    _check_all_modes(u, u.atoms[:900], u.atoms[900:], nbins=30, rng=(3.0, 12.4))
    # flat triclinic cell with a large offset along one lattice direction
    dims = [90.0, 35.0, 30.0, 70.0, 95.0, 100.0]
    pos = synthetic.uniform_triclinic(2, 4000, dims, rng)
    u = Universe(pos, dims)
    _check_all_modes(u, u.atoms, u.atoms, rng=(0.0, 11.0))


def test_sliced_tail_work_items_do_not_change_the_counts():
    """the last work items of a launch (here: all of them, the systems are
    small) are cut into slices of their j-clusters; debug flags bits 8-11
    force the slice count"""
    rng = np.random.default_rng(99)
    L = 36.0
    pos = (rng.random((5, 2100, 3)) * L).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    want, _, _ = oracle_interrdf(u, u.atoms, u.atoms, 60, (0.0, 13.0))
    want2, _, _ = oracle_interrdf(u, u.atoms[:300], u.atoms[300:], 60, (0.0, 13.0))
    dims = [30.0, 31.0, 32.0, 70.0, 80.0, 65.0]
    ut = Universe(synthetic.uniform_triclinic(3, 1500, dims, rng), dims)
    want3, _, _ = oracle_interrdf(ut, ut.atoms, ut.atoms, 50, (0.0, 11.0))
    for ns in (0, 1, 2, 3, 5, 8, 15):
        flags = ns << 8
        np.testing.assert_array_equal(
            _gpu(u, u.atoms, u.atoms, 60, (0.0, 13.0), debug=flags), want)
        np.testing.assert_array_equal(
            _gpu(u, u.atoms[:300], u.atoms[300:], 60, (0.0, 13.0), debug=flags),
            want2)
        np.testing.assert_array_equal(
            _gpu(ut, ut.atoms, ut.atoms, 50, (0.0, 11.0), debug=flags | 3), want3)
        np.testing.assert_array_equal(
            _gpu(ut, ut.atoms, ut.atoms, 50, (0.0, 11.0), debug=flags), want3)


def _last_stats():
    import torch
    from pmda_b200 import engine
    st = engine._state(torch.device("cuda", torch.cuda.current_device()))
    ws = st.workspaces[0]
    ws = ws[(-ws.data_ptr()) % 256:]
    return _cabi.rdf_stats(ws, st.compute_streams[0].cuda_stream)


def test_bins_narrower_than_the_error_bound_take_the_fp64_path():
    """a frame whose fp32 error bound is not far below the bin width cannot use
    the drain's candidate bin (it must stay <= nbins): params_kernel sends it
    through the reference arithmetic, pair by pair"""
    rng = np.random.default_rng(118)
    L = 30.0
    far = np.array([30000.0, -30000.0, 30000.0])
    pos = (rng.random((2, 500, 3)) * L + far).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    # bound ~ 3e-3 A against dr = 12 / 2000 = 6e-3 A (shared-memory histogram:
    # the range-from-0 kernel)
    want, _, _ = oracle_interrdf(u, u.atoms, u.atoms, 2000, (0.0, 12.0), None)
    for debug in (None, 16):
        got = _gpu(u, u.atoms, u.atoms, nbins=2000, rng=(0.0, 12.0), debug=debug)
        np.testing.assert_array_equal(got, want, err_msg="debug flags %r" % debug)
        assert _last_stats()["all_slow_frames"] == 2
    # the same atoms near the origin keep the fast path
    pos = (pos - far.astype(np.float32)).astype(np.float32)
    u = Universe(pos, [L, L, L, 90, 90, 90])
    want, _, _ = oracle_interrdf(u, u.atoms, u.atoms, 2000, (0.0, 12.0), None)
    got = _gpu(u, u.atoms, u.atoms, nbins=2000, rng=(0.0, 12.0))
    np.testing.assert_array_equal(got, want)
    assert _last_stats()["all_slow_frames"] == 0


def test_repeated_small_calls_replay_a_graph_with_the_same_integers():
    """librdfk captures a small call that comes back with identical arguments
    into a CUDA graph and replays it: same integers, same launch accounting,
    new coordinates behind the same pointers are seen, debug flag 32 turns it
    off"""
    from pmda_b200 import engine
    from pmda_b200.universe import MemoryTrajectory
    rng = np.random.default_rng(119)
    L = 34.0
    pos = (rng.random((6, 1500, 3)) * L).astype(np.float32)
    u = Universe(pos.copy(), [L, L, L, 90, 90, 90])
    want, _, _ = oracle_interrdf(u, u.atoms, u.atoms, 60, (0.0, 11.0), None)
    engine.clear_cache()
    replays0 = _cabi.graph_replay_count()
    per_run = []
    for rep in range(5):
        l0 = _cabi.launch_count()
        got = _gpu(u, u.atoms, u.atoms, nbins=60, rng=(0.0, 11.0))
        per_run.append(_cabi.launch_count() - l0)
        np.testing.assert_array_equal(got, want, err_msg="run %d" % rep)
    assert _cabi.graph_replay_count() - replays0 >= 2, per_run
    assert len(set(per_run)) == 1, per_run
    # an in-place edit of the trajectory: the cached block is re-validated by
    # its hash, re-uploaded, and the replayed graph reads the new frames
    traj = u.trajectory
    assert isinstance(traj, MemoryTrajectory)
    traj.positions[3, 7, :] += np.float32(2.5)
    want2, _, _ = oracle_interrdf(u, u.atoms, u.atoms, 60, (0.0, 11.0), None)
    assert (want2 != want).any()
    for rep in range(3):
        got = _gpu(u, u.atoms, u.atoms, nbins=60, rng=(0.0, 11.0))
        np.testing.assert_array_equal(got, want2, err_msg="edited, run %d" % rep)
    # switched off: no further replays
    r1 = _cabi.graph_replay_count()
    for rep in range(3):
        got = _gpu(u, u.atoms, u.atoms, nbins=60, rng=(0.0, 11.0), debug=32)
        np.testing.assert_array_equal(got, want2)
    assert _cabi.graph_replay_count() == r1
