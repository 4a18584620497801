"""HydrogenBondAnalysis.

CPU part: the oracle's restatement of capped_distance(min_cutoff) + calc_angles
against independent float64 geometry, and the host logic of
pmda_b200.hbond_analysis (block partition, D-H pairing, update_selections,
summaries) with the kernels replaced by tests/fake_backend.py.  Mirrors the
structure of pmda/test/test_hbond_analysis.py (count_by_time / count_by_ids,
start-stop-step, no-bond systems); its golden numbers need MDAnalysisTests
data files that are absent here.

GPU part (`-m gpu`): rdfk_hbonds through the public API against the oracle:
the hydrogen-bond set and the distances bit for bit, the angles to 1e-11
degrees (CUDA atan2 vs glibc atan2)."""
import numpy as np
import pytest

from pmda_b200 import engine, synthetic
from pmda_b200.hbond_analysis import HydrogenBondAnalysis
from pmda_b200.universe import Universe

from oracle import c_oracle, numpy_oracle
from fake_backend import OracleBackend

ANGLE_TOL = 1e-11     # degrees


def water_universe(n_frames=4, n_mol=150, dims=(17.0, 17.0, 17.0, 90, 90, 90),
                   seed=5):
    rng = np.random.default_rng(seed)
    if dims is None:
        pos, names = synthetic.water_box(n_frames, n_mol,
                                         (17.0, 17.0, 17.0, 90, 90, 90), rng)
        return Universe(pos, None, names=names)
    pos, names = synthetic.water_box(n_frames, n_mol, dims, rng)
    return Universe(pos, list(dims), names=names)


def water_pairs(u):
    """the bonded D-H pairs of the O H H layout"""
    n = u.atoms.n_atoms // 3
    o = 3 * np.arange(n)
    return np.repeat(o, 2), (o[:, None] + np.array([1, 2])).reshape(-1)


def oracle_hbonds(u, donors, hydrogens, acceptors, frames, d_a=3.0, ang=150.0,
                  fast=False):
    rows = []
    fn = c_oracle.distance_array if fast else None
    for i in frames:
        ts = u.trajectory[i]
        rows.append(numpy_oracle.hbonds_single_frame(
            ts.frame, ts.positions, donors, hydrogens, acceptors,
            ts.dimensions, d_a, ang, distance_fn=fn))
    return np.vstack(rows) if rows else np.empty((0, 6))


def assert_same_hbonds(got, want):
    assert got.shape == want.shape
    np.testing.assert_array_equal(got[:, :5], want[:, :5])   # ids + distance
    np.testing.assert_allclose(got[:, 5], want[:, 5], rtol=0, atol=ANGLE_TOL)


# ---------------------------------------------------------------- oracle KATs
def test_calc_angles_known_answers():
    a = np.array([[1, 0, 0], [1, 0, 0], [1, 0, 0], [0, 2, 0]], np.float32)
    b = np.zeros((4, 3), np.float32)
    c = np.array([[0, 1, 0], [-1, 0, 0], [1, 1, 0], [0, 5, 0]], np.float32)
    got = np.rad2deg(numpy_oracle.calc_angles(a, b, c))
    np.testing.assert_allclose(got, [90.0, 180.0, 45.0, 0.0], atol=1e-12)
    # across a periodic boundary: apex at the cell edge, arms in the next image
    box = [10.0, 10.0, 10.0, 90, 90, 90]
    a = np.array([[9.5, 5.0, 5.0]], np.float32)
    b = np.array([[0.5, 5.0, 5.0]], np.float32)     # 1 A from a through the wall
    c = np.array([[1.5, 5.0, 5.0]], np.float32)
    assert np.rad2deg(numpy_oracle.calc_angles(a, b, c, box))[0] == \
        pytest.approx(180.0, abs=1e-5)
    assert np.rad2deg(numpy_oracle.calc_angles(a, b, c))[0] < 1e-5


@pytest.mark.parametrize("dims", [
    [21.0, 23.0, 25.0, 90, 90, 90], [24.0, 25.0, 26.0, 75.0, 80.0, 70.0], None])
def test_oracle_against_independent_float64_geometry(dims):
    """125-image float64 search (not the reference's algorithm): same pair set
    and angles away from the decision boundaries"""
    rng = np.random.default_rng(9)
    u = water_universe(1, 90, tuple(dims) if dims else None, seed=9)
    pos = u.trajectory[0].positions.astype(np.float64)
    don, hyd = water_pairs(u)
    acc = 3 * np.arange(90)
    got = oracle_hbonds(u, don, hyd, acc, [0])
    if dims is None:
        shifts = np.zeros((1, 3))
    else:
        m = np.asarray(numpy_oracle.triclinic_vectors(dims), np.float64)
        g = np.arange(-2, 3)
        shifts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) @ m

    def mi(v):
        cand = v[None, :] + shifts
        return cand[np.argmin((cand * cand).sum(1))]

    want = []
    for k in range(len(don)):
        for j in range(len(acc)):
            v = mi(pos[acc[j]] - pos[don[k]])
            d = np.sqrt(v @ v)
            if not (1.0 + 1e-6 < d < 3.0 - 1e-6):
                continue
            r1, r2 = mi(pos[don[k]] - pos[hyd[k]]), mi(pos[acc[j]] - pos[hyd[k]])
            cosang = r1 @ r2 / np.sqrt((r1 @ r1) * (r2 @ r2))
            ang = np.rad2deg(np.arccos(np.clip(cosang, -1, 1)))
            if ang > 150.0 + 1e-6:
                want.append((don[k], hyd[k], acc[j], d, ang))
    want = np.array(want).reshape(-1, 5)
    # drop oracle rows sitting on a boundary the float64 check excluded
    keep = (np.abs(got[:, 4] - 3.0) > 1e-5) & (np.abs(got[:, 4] - 1.0) > 1e-5) \
        & (np.abs(got[:, 5] - 150.0) > 1e-5)
    got = got[keep]
    assert len(want) > 3
    np.testing.assert_array_equal(got[:, 1:4], want[:, :3])
    np.testing.assert_allclose(got[:, 4], want[:, 3], rtol=1e-6)
    np.testing.assert_allclose(got[:, 5], want[:, 4], atol=1e-3)


def test_capped_distance_min_cutoff_is_exclusive_max_inclusive():
    ref = np.zeros((1, 3), np.float32)
    conf = np.array([[1.0, 0, 0], [3.0, 0, 0], [2.0, 0, 0], [0.5, 0, 0]],
                    np.float32)
    pairs, d = numpy_oracle.capped_distance_minmax(ref, conf, 3.0, 1.0)
    assert pairs[:, 1].tolist() == [1, 2] and d.tolist() == [3.0, 2.0]
    pairs, d = numpy_oracle.capped_distance_minmax(ref, conf, 3.0, None)
    assert pairs[:, 1].tolist() == [0, 1, 2, 3]


# ------------------------------------------------------------ golden fixtures
import hashlib                                           # noqa: E402
import os                                                # noqa: E402
import sys                                               # noqa: E402

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "golden"))
import make_hbond_golden as mkhb                         # noqa: E402

HB_GOLD = np.load(os.path.join(_HERE, "golden", "hbond_rows.npz"))


@pytest.mark.parametrize("name", sorted(mkhb.CASES))
def test_oracle_reproduces_golden_hbond_tables(name):
    rows, digest = mkhb.oracle_rows(name)
    assert digest == HB_GOLD[name + "__sha256"].tobytes().decode()
    want = HB_GOLD[name + "__rows"]
    assert len(want) > 10
    np.testing.assert_array_equal(rows[:, :5], want[:, :5])
    np.testing.assert_allclose(rows[:, 5], want[:, 5], rtol=0, atol=ANGLE_TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mkhb.CASES))
def test_cuda_path_matches_golden_hbond_tables(name):
    pos, dims, names, don, hyd, acc, d_a, ang = mkhb.make_inputs(name)
    assert hashlib.sha256(pos.tobytes()).hexdigest() == \
        HB_GOLD[name + "__sha256"].tobytes().decode()
    u = Universe(pos, dims, names=names)
    h = HydrogenBondAnalysis(u, acceptors_sel=acc, pairs=(don, hyd),
                             d_a_cutoff=d_a, d_h_a_angle_cutoff=ang).run(
        n_blocks=min(2, len(pos)))
    assert_same_hbonds(h.hbonds, HB_GOLD[name + "__rows"])


# ------------------------------------------------------------------ host logic
@pytest.fixture
def backend():
    be = OracleBackend(n_devices=2)
    engine._set_test_backend(be)
    engine.clear_cache()
    yield be
    engine._set_test_backend(None)
    engine.clear_cache()


@pytest.mark.parametrize("dims", [
    (17.0, 17.0, 17.0, 90, 90, 90), (18.0, 19.0, 20.0, 80.0, 75.0, 70.0), None])
@pytest.mark.parametrize("n_blocks,n_jobs", [(1, 1), (3, 2)])
def test_host_matches_oracle_loop(backend, dims, n_blocks, n_jobs):
    u = water_universe(5, 60, dims)
    don, hyd = water_pairs(u)
    acc = u.select_atoms("name OW").indices
    h = HydrogenBondAnalysis(u, acceptors_sel="name OW", pairs=(don, hyd))
    h.run(n_blocks=n_blocks, n_jobs=n_jobs)
    want = oracle_hbonds(u, don, hyd, acc, range(5))
    assert len(want) > 5
    np.testing.assert_array_equal(h.hbonds, want)
    assert backend.calls > 0
    np.testing.assert_array_equal(
        h.count_by_time(),
        [np.sum(want[:, 0] == f) for f in range(5)])
    ids = h.count_by_ids()
    assert ids[:, 3].sum() == len(want)
    assert np.all(np.diff(ids[:, 3]) <= 0)


def oracle_dh_pairs(u, frame, don_sel="name OW", hyd_sel="name HW1 HW2"):
    """_get_dh_pairs through d_h_cutoff (pmda/hbond_analysis.py:366-381)"""
    ts = u.trajectory[frame]
    d = u.select_atoms(don_sel).indices
    hy = u.select_atoms(hyd_sel).indices
    pairs, _ = numpy_oracle.capped_distance(ts.positions[d], ts.positions[hy],
                                            1.2, box=ts.dimensions)
    return d[pairs[:, 0]], hy[pairs[:, 1]]


def test_d_h_cutoff_pairing_and_update_selections(backend):
    u = water_universe(4, 50)
    acc = u.select_atoms("name OW").indices
    for upd in (True, False):
        u.trajectory[0]
        h = HydrogenBondAnalysis(u, donors_sel="name OW",
                                 hydrogens_sel="name HW1 HW2",
                                 acceptors_sel="name OW",
                                 update_selections=upd).run(n_blocks=2)
        want = np.vstack([
            oracle_hbonds(u, *oracle_dh_pairs(u, i if upd else 0), acc, [i])
            for i in range(4)])
        assert len(want) > 5
        np.testing.assert_array_equal(h.hbonds, want)


def test_update_selections_follows_changing_pairs(backend):
    """a hydrogen that leaves its donor stops being paired in those frames"""
    u = water_universe(4, 40, seed=12)
    pos = u.trajectory.positions
    pos[2:, 1] += np.float32(3.0)      # H1 of molecule 0 flies off in frames 2, 3
    u = Universe(pos, u.trajectory.dimensions, names=u.names)
    don_sel, hyd_sel = "name OW", "name HW1 HW2"
    acc = u.select_atoms("name OW").indices
    h = HydrogenBondAnalysis(u, don_sel, hyd_sel, "name OW").run()
    rows = []
    for i in range(4):
        ts = u.trajectory[i]
        d = u.select_atoms(don_sel).indices
        hy = u.select_atoms(hyd_sel).indices
        pairs, _ = numpy_oracle.capped_distance(ts.positions[d],
                                                ts.positions[hy], 1.2,
                                                box=ts.dimensions)
        rows.append(numpy_oracle.hbonds_single_frame(
            i, ts.positions, d[pairs[:, 0]], hy[pairs[:, 1]], acc,
            ts.dimensions))
    np.testing.assert_array_equal(h.hbonds, np.vstack(rows))
    # frozen selections keep the frame-0 pairs
    u.trajectory[0]
    h2 = HydrogenBondAnalysis(u, don_sel, hyd_sel, "name OW",
                              update_selections=False).run()
    d = u.select_atoms(don_sel).indices
    hy = u.select_atoms(hyd_sel).indices
    ts0 = u.trajectory[0]
    pairs, _ = numpy_oracle.capped_distance(ts0.positions[d], ts0.positions[hy],
                                            1.2, box=ts0.dimensions)
    want = oracle_hbonds(u, d[pairs[:, 0]], hy[pairs[:, 1]], acc, range(4))
    np.testing.assert_array_equal(h2.hbonds, want)
    # ... wherever the caller's reader stands: the reference pairs on a fresh
    # Universe at frame 0 (pmda/hbond_analysis.py:398-408)
    u.trajectory[3]
    calls = backend.calls
    h3 = HydrogenBondAnalysis(u, don_sel, hyd_sel, "name OW",
                              update_selections=False).run()
    np.testing.assert_array_equal(h3.hbonds, want)
    # with update_selections the pairs of every frame come from one search
    # over the run: no extra pairing search in _prepare
    per_run = backend.calls - calls
    calls = backend.calls
    HydrogenBondAnalysis(u, don_sel, hyd_sel, "name OW").run()
    assert backend.calls - calls <= per_run + 4


def test_start_stop_step_and_single_frame(backend):
    u = water_universe(9, 40)
    don, hyd = water_pairs(u)
    acc = u.select_atoms("name OW").indices
    h = HydrogenBondAnalysis(u, acceptors_sel=acc, pairs=(don, hyd))
    h.run(start=1, stop=8, step=3, n_blocks=2)
    want = oracle_hbonds(u, don, hyd, acc, range(1, 8, 3))
    np.testing.assert_array_equal(h.hbonds, want)
    np.testing.assert_array_equal(h.frames, [1, 4, 7])
    np.testing.assert_array_equal(
        h.count_by_time(), [np.sum(want[:, 0] == f) for f in (1, 4, 7)])
    one = h._single_frame(u.trajectory[4], None)
    np.testing.assert_array_equal(one, want[want[:, 0] == 4])


def test_no_bond_and_no_hydrogen_systems(backend):
    u = water_universe(3, 30)
    don, hyd = water_pairs(u)
    far = HydrogenBondAnalysis(u, acceptors_sel="name OW", pairs=(don, hyd),
                               d_h_a_angle_cutoff=179.999).run()
    assert far.hbonds.shape == (0, 6)
    assert far.count_by_time().tolist() == [0, 0, 0]
    with pytest.raises(ValueError):
        HydrogenBondAnalysis(u, hydrogens_sel="name HW1",
                             acceptors_sel="name OW").run()
    with pytest.raises(ValueError):
        HydrogenBondAnalysis(u, pairs=(don, hyd)).run()       # no acceptors
    with pytest.raises(ValueError):
        HydrogenBondAnalysis(u, pairs=(don, hyd[:-1]))
    with pytest.raises(NotImplementedError):
        HydrogenBondAnalysis(u).guess_acceptors()


# ------------------------------------------------------------------ GPU parity
def _gpu_run(u, **kw):
    """runs twice: through the acceptor cell grid and, with debug flag 1,
    through the brute-force kernel; both must give the same table"""
    from pmda_b200 import _cabi
    run_kw = {k: kw.pop(k) for k in ("n_blocks", "n_jobs", "start", "stop",
                                     "step") if k in kw}
    old = _cabi.set_debug_flags(0)
    try:
        h = HydrogenBondAnalysis(u, **kw).run(**run_kw)
        _cabi.set_debug_flags(1)
        brute = HydrogenBondAnalysis(u, **kw).run(**run_kw)
    finally:
        _cabi.set_debug_flags(old)
    np.testing.assert_array_equal(h.hbonds, brute.hbonds)
    return h


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [
    (17.0, 17.0, 17.0, 90, 90, 90), (25.0, 19.0, 31.0, 90, 90, 90),
    (18.0, 19.0, 20.0, 80.0, 75.0, 70.0), (20.0, 20.0, 20.0, 60.0, 60.0, 90.0),
    None])
def test_gpu_matches_oracle(dims):
    u = water_universe(5, 260, dims, seed=21)
    don, hyd = water_pairs(u)
    acc = u.select_atoms("name OW").indices
    h = _gpu_run(u, acceptors_sel="name OW", pairs=(don, hyd), n_blocks=2)
    want = oracle_hbonds(u, don, hyd, acc, range(5))
    assert len(want) > 20
    assert_same_hbonds(h.hbonds, want)
    # a loose angle and a long cut-off: many more candidates per donor
    h = _gpu_run(u, acceptors_sel="name OW", pairs=(don, hyd), d_a_cutoff=4.5,
                 d_h_a_angle_cutoff=100.0)
    want = oracle_hbonds(u, don, hyd, acc, range(5), d_a=4.5, ang=100.0)
    assert_same_hbonds(h.hbonds, want)


@pytest.mark.gpu
def test_gpu_config1_sized_water_box():
    """3,000 waters (BASELINE config 1's system with its hydrogens): 6,000 D-H
    rows x 3,000 acceptors per frame, oracle through its C distance matrix"""
    L = 44.78
    u = water_universe(3, 3000, (L, L, L, 90, 90, 90), seed=22)
    don, hyd = water_pairs(u)
    acc = u.select_atoms("name OW").indices
    h = _gpu_run(u, acceptors_sel="name OW", pairs=(don, hyd), n_blocks=2)
    want = oracle_hbonds(u, don, hyd, acc, range(3), fast=True)
    assert len(want) > 100
    assert_same_hbonds(h.hbonds, want)


@pytest.mark.gpu
def test_gpu_d_h_cutoff_pairing_every_frame():
    """donors_sel route: D-H pairs from capped_distance(d_h_cutoff) on the GPU,
    re-derived per frame (update_selections) or frozen at the current frame"""
    u = water_universe(4, 300, (21.0, 21.0, 21.0, 90, 90, 90), seed=27)
    acc = u.select_atoms("name OW").indices
    for upd in (True, False):
        u.trajectory[0]
        h = _gpu_run(u, donors_sel="name OW", hydrogens_sel="name HW1 HW2",
                     acceptors_sel="name OW", update_selections=upd)
        want = np.vstack([
            oracle_hbonds(u, *oracle_dh_pairs(u, i if upd else 0), acc, [i])
            for i in range(4)])
        assert len(want) > 20
        assert_same_hbonds(h.hbonds, want)


@pytest.mark.gpu
def test_gpu_fp64_only_frames_and_nonfinite_coordinates():
    # cut-off beyond the triclinic brick: every pair on the fp64 path
    dims = (9.0, 9.5, 10.0, 70.0, 80.0, 65.0)
    u = water_universe(2, 70, dims, seed=23)
    don, hyd = water_pairs(u)
    acc = u.select_atoms("name OW").indices
    h = _gpu_run(u, acceptors_sel=acc, pairs=(don, hyd), d_a_cutoff=6.0,
                 d_h_a_angle_cutoff=120.0)
    want = oracle_hbonds(u, don, hyd, acc, range(2), d_a=6.0, ang=120.0)
    assert len(want) > 10
    assert_same_hbonds(h.hbonds, want)
    # NaN / Inf atoms: their pairs are never within the cut-off
    u = water_universe(2, 120, seed=24)
    pos = u.trajectory.positions.copy()
    pos[0, 3, 1] = np.nan
    pos[1, 30, 0] = np.inf
    u = Universe(pos, u.trajectory.dimensions, names=u.names)
    h = _gpu_run(u, acceptors_sel=acc[:120], pairs=water_pairs(u))
    want = oracle_hbonds(u, *water_pairs(u), acc[:120], range(2))
    assert_same_hbonds(h.hbonds, want)


@pytest.mark.gpu
def test_gpu_far_unwrapped_molecules_and_frame_varying_box():
    rng = np.random.default_rng(25)
    F, n_mol = 4, 200
    Ls = 18.0 + rng.random((F, 3)) * 2.0
    pos, names = synthetic.water_box(F, n_mol, (18.0, 18.0, 18.0, 90, 90, 90),
                                     rng)
    pos[1, :90] += np.float32(2 * 18.5)      # whole molecules images away
    pos[2, 300:330] -= np.float32(5 * 19.0)
    dims = np.concatenate([Ls, np.full((F, 3), 90.0)], 1).astype(np.float32)
    u = Universe(pos, dims, names=names)
    don, hyd = water_pairs(u)
    acc = u.select_atoms("name OW").indices
    h = _gpu_run(u, acceptors_sel=acc, pairs=(don, hyd))
    want = oracle_hbonds(u, don, hyd, acc, range(F))
    assert_same_hbonds(h.hbonds, want)


@pytest.mark.gpu
def test_gpu_cell_grid_shapes():
    """grids the 27-cell walk has to get right: one periodic axis off, a thin
    slab (a single cell across), a long box (more cells than the per-axis
    cap), a box barely wider than two cut-offs (periodic axis falls back to one
    cell)"""
    rng = np.random.default_rng(28)
    for dims, n_mol in (((30.0, 30.0, 0.0, 90, 90, 90), 500),
                        ((40.0, 40.0, 3.5, 90, 90, 90), 300),
                        ((150.0, 12.0, 12.0, 90, 90, 90), 600),
                        ((6.5, 6.5, 6.5, 90, 90, 90), 12),
                        ((9.5, 30.0, 30.0, 90, 90, 90), 300)):
        box = tuple(v if v > 0 else 25.0 for v in dims[:3]) + (90, 90, 90)
        pos, names = synthetic.water_box(3, n_mol, box, rng)
        u = Universe(pos, list(dims), names=names)
        don, hyd = water_pairs(u)
        acc = u.select_atoms("name OW").indices
        h = _gpu_run(u, acceptors_sel=acc, pairs=(don, hyd))
        want = oracle_hbonds(u, don, hyd, acc, range(3))
        assert len(want) > 0, dims
        assert_same_hbonds(h.hbonds, want)
    # clustered, no box: the grid spans the extents
    pos, names = synthetic.water_box(2, 400, (20.0, 20.0, 20.0, 90, 90, 90), rng)
    pos[:, 600:] += np.float32(35.0)
    u = Universe(pos, None, names=names)
    don, hyd = water_pairs(u)
    acc = u.select_atoms("name OW").indices
    h = _gpu_run(u, acceptors_sel=acc, pairs=(don, hyd))
    assert_same_hbonds(h.hbonds, oracle_hbonds(u, don, hyd, acc, range(2)))


@pytest.mark.gpu
def test_gpu_capped_pairs_and_ragged_capacity_through_the_cabi():
    """rdfk_hbonds called directly: plain capped_distance mode, and a result
    that does not fit `capacity` (total reported, first rows written)"""
    import torch
    from pmda_b200 import _cabi
    u = water_universe(3, 150, seed=26)
    pos = u.trajectory.positions
    dev = torch.device("cuda", 0)
    xyz = torch.as_tensor(pos, device=dev)
    o = torch.arange(0, 450, 3, dtype=torch.int32, device=dev)
    hs = torch.as_tensor(np.setdiff1d(np.arange(450), np.arange(0, 450, 3)),
                         dtype=torch.int32, device=dev)
    box = torch.zeros((3, 9), dtype=torch.float32, device=dev)
    box[:, :3] = 17.0
    ws = torch.empty(_cabi.hbonds_workspace_bytes(3, 150, 300) + 256,
                     dtype=torch.uint8, device=dev)
    ws = ws[(-ws.data_ptr()) % 256:]
    total = torch.zeros(1, dtype=torch.int64, device=dev)

    def call(cap):
        row = torch.full((max(cap, 1),), -7, dtype=torch.int32, device=dev)
        acc = torch.full((max(cap, 1),), -7, dtype=torch.int32, device=dev)
        dist = torch.zeros(max(cap, 1), dtype=torch.float64, device=dev)
        _cabi.hbonds(xyz, 3, 450, o, None, 150, hs, 300, _cabi.BOX_ORTHO, box,
                     1.2, None, 0.0, cap, row, acc, dist, None, total, ws,
                     torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        return int(total[0]), row.cpu().numpy(), acc.cpu().numpy(), \
            dist.cpu().numpy()

    want_rows, want_acc, want_d = [], [], []
    hs_np = hs.cpu().numpy()
    for f in range(3):
        pairs, d = numpy_oracle.capped_distance(
            pos[f][::3], pos[f][hs_np], 1.2, box=[17.0, 17.0, 17.0, 90, 90, 90])
        want_rows += list(f * 150 + pairs[:, 0])
        want_acc += list(pairs[:, 1])
        want_d += list(d)
    n, row, acc, dist = call(4096)
    assert n == len(want_rows) >= 900          # every O finds its two H
    np.testing.assert_array_equal(row[:n], want_rows)
    np.testing.assert_array_equal(acc[:n], want_acc)
    np.testing.assert_array_equal(dist[:n], want_d)
    n2, row, acc, dist = call(100)
    assert n2 == n
    np.testing.assert_array_equal(row[:100], want_rows[:100])
    np.testing.assert_array_equal(acc[:100], want_acc[:100])
    n3, row, _, _ = call(0)
    assert n3 == n and row[0] == -7
