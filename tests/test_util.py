"""Block partitioner + timer: same tables and properties as the reference's
pmda/test/test_util.py:23-130, plus golden outputs produced by the reference's
own make_balanced_slices (tests/golden/make_golden.py)."""
import json
import os
import time

import numpy as np
import pytest
from numpy.testing import assert_almost_equal, assert_equal

from pmda_b200.util import make_balanced_slices, timeit

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_timeit():
    with timeit() as timer:
        time.sleep(0.25)
    assert_almost_equal(timer.elapsed, 0.25, decimal=1)


def test_timeit_propagates_exceptions():
    with pytest.raises(KeyError):
        with timeit():
            raise KeyError("x")


@pytest.mark.parametrize("start", (None, 0, 1, 10))
@pytest.mark.parametrize("n_frames,n_blocks,result", [
    (5, 1, [slice(0, None, 1)]),
    (5, 2, [slice(0, 3, 1), slice(3, None, 1)]),
    (5, 3, [slice(0, 2, 1), slice(2, 4, 1), slice(4, None, 1)]),
    (5, 4, [slice(0, 2, 1), slice(2, 3, 1), slice(3, 4, 1),
            slice(4, None, 1)]),
    (5, 5, [slice(0, 1, 1), slice(1, 2, 1), slice(2, 3, 1), slice(3, 4, 1),
            slice(4, None, 1)]),
    (10, 2, [slice(0, 5, 1), slice(5, None, 1)]),
    (10, 3, [slice(0, 4, 1), slice(4, 7, 1), slice(7, None, 1)]),
    (10, 7, [slice(0, 2, 1), slice(2, 4, 1), slice(4, 6, 1), slice(6, 7, 1),
             slice(7, 8, 1), slice(8, 9, 1), slice(9, None, 1)]),
])
def test_make_balanced_slices_step1(n_frames, n_blocks, start, result):
    off = start if start is not None else 0
    want = [slice(sl.start + off,
                  sl.stop + off if sl.stop is not None else None, sl.step)
            for sl in result]
    got = make_balanced_slices(n_frames, n_blocks, start=start, step=1)
    assert_equal(got, want)


def _check_partition(n_blocks, start, stop, step, scale):
    traj_frames = range(scale * stop)
    frames = traj_frames[start:stop:step]
    n_frames = len(frames)
    if n_frames >= n_blocks:
        slices = make_balanced_slices(n_frames, n_blocks,
                                      start=start, stop=stop, step=step)
        assert len(slices) == n_blocks
        block_frames, block_sizes = [], []
        for bslice in slices:
            bframes = traj_frames[bslice]
            block_frames.extend(list(bframes))
            block_sizes.append(len(bframes))
        block_sizes = np.array(block_sizes)
        assert_equal(np.asarray(block_frames), np.asarray(frames))
        assert np.all(block_sizes > 0)
        minsize = n_frames // n_blocks
        assert len(np.setdiff1d(block_sizes, [minsize, minsize + 1])) == 0
    else:
        with pytest.raises(ValueError, match="n_blocks must be smaller"):
            make_balanced_slices(n_frames, n_blocks,
                                 start=start, stop=stop, step=step)


@pytest.mark.parametrize('n_blocks', [1, 2, 3, 4, 5, 7, 10, 11])
@pytest.mark.parametrize('start', [0, 1, 10])
@pytest.mark.parametrize('stop', [11, 100, 256])
@pytest.mark.parametrize('step', [None, 1, 2, 3, 5, 7])
@pytest.mark.parametrize('scale', [1, 2])
def test_make_balanced_slices(n_blocks, start, stop, step, scale):
    _check_partition(n_blocks, start, stop, step, scale)


def test_make_balanced_slices_step_gt_stop():
    _check_partition(n_blocks=2, start=None, stop=5, step=6, scale=1)


@pytest.mark.parametrize('n_blocks', [1, 2])
@pytest.mark.parametrize('start', [0, 10])
@pytest.mark.parametrize('step', [None, 1, 2])
def test_make_balanced_slices_empty(n_blocks, start, step):
    assert make_balanced_slices(0, n_blocks, start=start, step=step) == []


@pytest.mark.parametrize("n_frames,n_blocks,start,stop,step",
                         [(-1, 5, None, None, None), (5, 0, None, None, None),
                          (5, -1, None, None, None), (0, 0, None, None, None),
                          (-1, -1, None, None, None),
                          (5, 4, -1, None, None), (0, 5, -1, None, None),
                          (5, 0, -1, None, None),
                          (5, 4, None, -1, None), (5, 4, 3, 2, None),
                          (5, 4, None, None, -1), (5, 4, None, None, 0),
                          (4, 5, None, None, None)])
def test_make_balanced_slices_ValueError(n_frames, n_blocks, start, stop,
                                         step):
    with pytest.raises(ValueError):
        make_balanced_slices(n_frames, n_blocks,
                             start=start, stop=stop, step=step)


def test_golden_against_reference_implementation():
    """567 cases computed by /root/reference/pmda/util.py itself"""
    with open(os.path.join(GOLDEN, "balanced_slices.json")) as fh:
        cases = json.load(fh)
    assert len(cases) > 500
    for c in cases:
        kw = dict(start=c["start"], stop=c["stop"], step=c["step"])
        if "error" in c:
            with pytest.raises(ValueError):
                make_balanced_slices(c["n_frames"], c["n_blocks"], **kw)
        else:
            got = make_balanced_slices(c["n_frames"], c["n_blocks"], **kw)
            got = [[None if v is None else int(v)
                    for v in (s.start, s.stop, s.step)] for s in got]
            assert got == c["slices"], c
