"""Block partitioner and wall-clock timer (mirror of pmda/util.py:26-186).

``make_balanced_slices`` is the one scheduling rule of the reference: it decides
which frames form a block (pmda/parallel.py:365).  The GPU scheduler reuses it
for frames -> blocks and for blocks -> ranks, so ``_blocks`` and ``_results``
keep the reference's meaning for any GPU count.
"""
import time


class timeit(object):
    """Context manager recording wall-clock seconds in ``elapsed``
    (pmda/util.py:26-61)."""

    def __enter__(self):
        self._t0 = time.time()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.elapsed = time.time() - self._t0
        return False  # never swallow exceptions


def make_balanced_slices(n_frames, n_blocks, start=None, stop=None, step=None):
    """Split ``n_frames`` (already sliced by start:stop:step) into ``n_blocks``
    contiguous slices whose sizes differ by at most one frame and are never
    zero: the first ``n_frames % n_blocks`` blocks get the extra frame.

    Same contract as pmda/util.py:64-186: ``[]`` for ``n_frames == 0``;
    ``ValueError`` for negative sizes, ``n_blocks < 1``, ``n_blocks > n_frames``,
    ``start < 0``, ``stop < start`` or ``step < 1``; the last slice stops at
    ``stop`` (or ``None`` when no stop was given).
    """
    start = 0 if start is None else int(start)
    stop = None if stop is None else int(stop)
    step = 1 if step is None else int(step)

    if n_frames < 0:
        raise ValueError("n_frames must be >= 0")
    if n_blocks < 1:
        raise ValueError("n_blocks must be > 0")
    if n_frames != 0 and n_blocks > n_frames:
        raise ValueError("n_blocks must be smaller than n_frames: "
                         "{}".format(n_frames))
    if start < 0:
        raise ValueError("start must be >= 0 or None")
    if stop is not None and stop < start:
        raise ValueError("stop must be >= start and >= 0 or None")
    if step < 1:
        raise ValueError("step must be > 0 or None")

    if n_frames == 0:
        return []

    n_frames, n_blocks = int(n_frames), int(n_blocks)
    base, extra = divmod(n_frames, n_blocks)
    slices = []
    first = start
    for b in range(n_blocks):
        size = base + (1 if b < extra else 0)
        last = first + size * step
        slices.append(slice(first, last, step))
        first = last
    # the final boundary may overshoot the trajectory by < step; clamp it to
    # the caller's stop, or leave it open when there is none
    tail = slices[-1]
    tail_stop = None if stop is None else min(tail.stop, stop)
    slices[-1] = slice(tail.start, tail_stop, tail.step)
    return slices
