"""Operator protocol / scheduler: the reference's pmda/test/test_parallel.py
re-expressed on an in-memory trajectory (no MDAnalysis data files needed)."""
import numpy as np
import pytest

from pmda_b200 import parallel
from pmda_b200.universe import Universe
from pmda_b200.util import make_balanced_slices

N_FRAMES = 98       # like the DCD file the reference tests use


def test_timeing():
    io = np.arange(5)
    compute = np.arange(5) + 1
    total = 5
    universe = np.arange(2)
    prepare = 3
    conclude = 6
    wait = 12
    io_block = np.sum(io)
    compute_block = np.sum(compute)

    timing = parallel.Timing(io, compute, total,
                             universe, prepare, conclude, wait,
                             io_block, compute_block,)

    np.testing.assert_equal(timing.io, io)
    np.testing.assert_equal(timing.compute, compute)
    np.testing.assert_equal(timing.total, total)
    np.testing.assert_equal(timing.universe, universe)
    np.testing.assert_equal(timing.cumulate_time, np.sum(io) + np.sum(compute))
    np.testing.assert_equal(timing.prepare, prepare)
    np.testing.assert_equal(timing.conclude, conclude)
    np.testing.assert_equal(timing.wait, wait)
    np.testing.assert_equal(timing.io_block, io_block)
    np.testing.assert_equal(timing.compute_block, compute_block)


class NoneAnalysis(parallel.ParallelAnalysisBase):
    def __init__(self, atomgroup):
        universe = atomgroup.universe
        super(NoneAnalysis, self).__init__(universe, (atomgroup, ))

    def _prepare(self):
        pass

    def _conclude(self):
        self.res = np.concatenate(self._results)

    def _single_frame(self, ts, atomgroups):
        return ts.frame


def make_universe():
    rng = np.random.default_rng(0)
    pos = rng.random((N_FRAMES, 20, 3)).astype(np.float32) * 10
    return Universe(pos, [10, 10, 10, 90, 90, 90])


@pytest.fixture
def analysis():
    return NoneAnalysis(make_universe().atoms)


@pytest.mark.parametrize('n_jobs', (1, 2))
def test_all_frames(analysis, n_jobs):
    analysis.run(n_jobs=n_jobs)
    assert len(analysis.res) == N_FRAMES
    np.testing.assert_equal(analysis.res, np.arange(N_FRAMES))


@pytest.mark.parametrize('n_jobs', (1, 2))
def test_sub_frames(analysis, n_jobs):
    analysis.run(start=10, stop=50, step=10, n_jobs=n_jobs)
    np.testing.assert_almost_equal(analysis.res, [10, 20, 30, 40])


@pytest.mark.parametrize('n_jobs', (1, 2, 3))
def test_no_frames(analysis, n_jobs):
    with pytest.warns(UserWarning):
        analysis.run(start=N_FRAMES, stop=N_FRAMES + 1, n_jobs=n_jobs)
    assert len(analysis.res) == 0
    np.testing.assert_equal(analysis.res, [])
    np.testing.assert_equal(analysis.timing.compute, [])
    np.testing.assert_equal(analysis.timing.io, [])
    np.testing.assert_equal(analysis.timing.io_block, [0])
    np.testing.assert_equal(analysis.timing.compute_block, [0])
    np.testing.assert_equal(analysis.timing.wait, [0])
    assert analysis.timing.universe == 0


def test_nframes_less_nblocks_warning(analysis):
    with pytest.raises(ValueError):
        with pytest.warns(UserWarning):
            analysis.run(stop=2, n_blocks=4, n_jobs=2)


@pytest.mark.parametrize('n_blocks', np.arange(1, 11))
def test_nblocks(analysis, n_blocks):
    analysis.run(n_blocks=n_blocks)
    assert len(analysis._results) == n_blocks


def test_guess_nblocks(analysis):
    # n_blocks defaults to n_jobs (pmda/parallel.py:335-337)
    analysis.run(n_jobs=3)
    assert len(analysis._results) == 3


@pytest.mark.parametrize('n_blocks', np.arange(1, 11))
def test_blocks(analysis, n_blocks):
    analysis.run(n_blocks=n_blocks)
    u = make_universe()
    start, stop, step = u.trajectory.check_slice_indices(None, None, None)
    slices = make_balanced_slices(N_FRAMES, n_blocks, start, stop, step)
    blocks = [range(b.start, b.stop, b.step) for b in slices]
    assert analysis._blocks == blocks


def test_timing_shapes(analysis):
    analysis.run(n_blocks=4)
    t = analysis.timing
    assert len(t.io) == N_FRAMES and len(t.compute) == N_FRAMES
    for per_block in (t.universe, t.wait, t.io_block, t.compute_block):
        assert len(per_block) == 4
    assert t.total >= 0 and t.prepare >= 0 and t.conclude >= 0


def test_attrlock():
    u = make_universe()
    pab = parallel.ParallelAnalysisBase(u, (u.atoms,))
    pab.thing1 = 24
    assert pab.thing1 == 24
    with pab.readonly_attributes():
        assert pab.thing1 == 24
        with pytest.raises(AttributeError):
            pab.thing2 = 100
    pab.thing2 = 100
    assert pab.thing2 == 100


def test_single_frame_not_implemented():
    u = make_universe()
    pab = parallel.ParallelAnalysisBase(u, (u.atoms,))
    with pytest.raises(NotImplementedError):
        pab.run(stop=1)


def test_reduce():
    res = []
    u = make_universe()
    ana = NoneAnalysis(u.atoms)
    res = ana._reduce(res, [1])
    res = ana._reduce(res, [1])
    assert res == [[1], [1]]


def test_check_slice_indices():
    traj = make_universe().trajectory
    assert traj.check_slice_indices(None, None, None) == (0, N_FRAMES, 1)
    assert traj.check_slice_indices(10, 50, 10) == (10, 50, 10)
    assert traj.check_slice_indices(-10, None, None) == (N_FRAMES - 10,
                                                         N_FRAMES, 1)
    assert traj.check_slice_indices(0, 10 ** 6, 2) == (0, N_FRAMES, 2)
    with pytest.raises(ValueError):
        traj.check_slice_indices(0, 10, 0)
    with pytest.raises(TypeError):
        traj.check_slice_indices(0.5, 10, 1)


def test_frame_cache_tokens_survive_id_reuse():
    """The device frame cache must never serve another trajectory's frames
    (ids are reused after GC)."""
    import gc
    from pmda_b200 import engine
    from pmda_b200.universe import Universe
    rng = np.random.default_rng(5)
    u1 = Universe(rng.random((2, 50, 3)).astype(np.float32), [9, 9, 9, 90, 90, 90])
    t1 = engine._traj_token(u1.trajectory)
    assert engine._traj_token(u1.trajectory) == t1
    del u1
    gc.collect()
    u2 = Universe(rng.random((2, 50, 3)).astype(np.float32), [9, 9, 9, 90, 90, 90])
    assert engine._traj_token(u2.trajectory) != t1
