"""Operator protocol + block scheduler (mirror of pmda/parallel.py:33-450).

Same public surface as the reference: subclasses implement ``_prepare``,
``_single_frame(ts, atomgroups)``, ``_reduce(res, frame_result)`` and
``_conclude``; users call ``run(start, stop, step, n_jobs, n_blocks)`` and read
``_results`` (one entry per block), ``_blocks``, ``timing``.

What changed underneath: the dask task graph of pmda/parallel.py:376-392 is
gone.  Blocks are still cut by ``make_balanced_slices`` (same ``_blocks``), but

* analyses that define ``_device_blocks`` (InterRDF, InterRDF_s) hand whole
  frame blocks to per-GPU streams (pmda_b200/engine.py) instead of looping
  ``_single_frame`` in worker processes;
* any other subclass runs the reference's per-frame loop
  (pmda/parallel.py:429-439) in-process, block after block -- that is user
  Python, there is nothing to put on a GPU.

``n_jobs`` keeps its meaning "how many workers": here, how many local GPUs the
run may use (-1 = all visible).  Under ``torch.distributed`` every rank calls
``run()`` with the same arguments and frames are sharded over ranks.
"""
from contextlib import contextmanager
import time
import warnings

import numpy as np

from .util import timeit, make_balanced_slices


def _tp(label):
    from . import engine
    engine._tp(label)


class Timing(object):
    """Timings gathered during ``run()`` (pmda/parallel.py:33-104).

    ``io`` / ``compute`` hold one value per frame, ``universe`` / ``wait`` /
    ``io_block`` / ``compute_block`` one per block.  On the GPU path ``io`` is
    the staging time (host gather + H2D) and ``compute`` the CUDA-event kernel
    time of the frame's batch, both divided evenly over the batch's frames.
    """

    def __init__(self, io, compute, total, universe, prepare,
                 conclude, wait=None, io_block=None,
                 compute_block=None):
        self._io = io
        self._io_block = io_block
        self._compute = compute
        self._compute_block = compute_block
        self._total = total
        self._cumulate = np.sum(io) + np.sum(compute)
        self._universe = universe
        self._prepare = prepare
        self._conclude = conclude
        self._wait = wait

    io = property(lambda self: self._io, doc="io time per frame")
    io_block = property(lambda self: self._io_block, doc="io time per block")
    compute = property(lambda self: self._compute,
                       doc="compute time per frame")
    compute_block = property(lambda self: self._compute_block,
                             doc="compute time per block")
    total = property(lambda self: self._total, doc="wall time")
    cumulate_time = property(lambda self: self._cumulate,
                             doc="sum of io and compute over all frames")
    universe = property(lambda self: self._universe,
                        doc="time to create a universe for each block")
    prepare = property(lambda self: self._prepare, doc="time to prepare")
    conclude = property(lambda self: self._conclude, doc="time to conclude")
    wait = property(lambda self: self._wait,
                    doc="time for blocks to start working")


def _stack_block_results(per_block):
    """``np.asarray(per_block)`` as pmda/parallel.py:399 does; blocks of unequal
    length (a ragged list, which numpy >= 1.24 refuses to guess) become a 1-D
    object array with one entry per block."""
    try:
        return np.asarray(per_block)
    except ValueError:
        out = np.empty(len(per_block), dtype=object)
        for i, r in enumerate(per_block):
            out[i] = r
        return out


def _visible_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:  # pragma: no cover - torch always present in this image
        return 0


class ParallelAnalysisBase(object):
    """Base class for multi-frame analyses split into frame blocks."""

    def __init__(self, universe, atomgroups):
        # same bookkeeping as pmda/parallel.py:210-213
        self._trajectory = universe.trajectory
        self._top = universe.filename
        self._traj = universe.trajectory.filename
        self._indices = [ag.indices for ag in atomgroups]
        # the reference re-opens the files in every worker; here workers are
        # GPU streams of this process, so the live objects are kept
        self._universe = universe
        self._atomgroups = tuple(atomgroups)

    @contextmanager
    def readonly_attributes(self):
        """Forbid attribute assignment while blocks are running
        (pmda/parallel.py:215-239)."""
        self._attr_lock = True
        try:
            yield
        finally:
            self._attr_lock = False

    def __setattr__(self, key, val):
        if key == '_attr_lock' or not getattr(self, '_attr_lock', False):
            super(ParallelAnalysisBase, self).__setattr__(key, val)
        else:
            raise AttributeError("Can't set attribute at this time")

    def _conclude(self):
        """Unpack ``self._results`` into user-facing attributes."""
        pass  # pylint: disable=unnecessary-pass

    def _prepare(self):
        """Extra set-up before the blocks start."""
        pass  # pylint: disable=unnecessary-pass

    def _single_frame(self, ts, atomgroups):
        """Per-frame computation; may only *read* ``self``."""
        raise NotImplementedError

    def run(self, start=None, stop=None, step=None, n_jobs=1, n_blocks=None):
        """Perform the calculation.

        start, stop, step : frame slice of the trajectory
        n_jobs   : number of workers (local GPUs); ``-1`` = all visible GPUs
        n_blocks : number of frame blocks (default ``n_jobs``)
        """
        if n_jobs == -1:
            n_jobs = max(1, _visible_gpus())
        if n_blocks is None:
            n_blocks = n_jobs

        start, stop, step = self._trajectory.check_slice_indices(start, stop,
                                                                 step)
        n_frames = len(range(start, stop, step))
        self.start, self.stop, self.step = start, stop, step
        self.n_frames = n_frames

        if n_frames == 0:
            warnings.warn("run() analyses no frames: check start/stop/step")
        if n_frames < n_blocks:
            warnings.warn("run() uses more blocks than frames: "
                          "decrease n_blocks")

        slices = make_balanced_slices(n_frames, n_blocks,
                                      start=start, stop=stop, step=step)

        _tp("run: slices")
        with timeit() as total:
            with timeit() as prepare:
                self._prepare()
            time_prepare = prepare.elapsed
            _blocks = [range(b.start, b.stop, b.step) for b in slices]
            with self.readonly_attributes():
                wait_start = time.time()
                device_blocks = getattr(self, '_device_blocks', None)
                if device_blocks is not None and len(_blocks):
                    _tp("run: prepare")
                    res = device_blocks(_blocks, n_jobs)
                    _tp("run: device blocks returned")
                else:
                    res = [self._block_helper(b) for b in _blocks]
            if len(res) == 0:
                # n_frames == 0: keep the shape the rest expects
                res = [([], [], [], 0, wait_start, 0, 0)]
            with timeit() as conclude:
                self._results = _stack_block_results([el[0] for el in res])
                self._blocks = _blocks
                self._conclude()
        _tp("run: conclude")
        self.timing = Timing(
            np.hstack([el[1] for el in res]),
            np.hstack([el[2] for el in res]), total.elapsed,
            np.array([el[3] for el in res]), time_prepare,
            conclude.elapsed,
            np.array([el[4] - wait_start for el in res]),
            np.array([el[5] for el in res]),
            np.array([el[6] for el in res]))
        _tp("run: timing object")
        return self

    def _block_helper(self, frames):
        """Generic per-frame loop of one block (pmda/parallel.py:415-444)."""
        wait_end = time.time()
        with timeit() as b_universe:
            agroups = self._atomgroups
        res = []
        times_io = []
        times_compute = []
        for i in frames:
            with timeit() as b_io:
                ts = self._trajectory[i]
            with timeit() as b_compute:
                res = self._reduce(res, self._single_frame(ts, agroups))
            times_io.append(b_io.elapsed)
            times_compute.append(b_compute.elapsed)
        return (np.asarray(res), np.asarray(times_io),
                np.asarray(times_compute), b_universe.elapsed, wait_end,
                np.sum(times_io), np.sum(times_compute))

    @staticmethod
    def _reduce(res, result_single_frame):
        """'append' action for a time series"""
        res.append(result_single_frame)
        return res
