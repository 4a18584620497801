"""The reference's CPU execution SHAPE, restated (test infrastructure, NOT
product code -- see the header of rdf_oracle.c for who may use oracle/).

pmda runs one *process* per frame block (dask `processes` / multiprocessing
scheduler, pmda/parallel.py:333-349, 392); each worker loops over its frames
(:429-439) and per frame executes InterRDF._single_frame (pmda/rdf.py:120-137):

    pairs, dist = capped_distance(g1.positions, g2.positions, rmax, box)
    count = np.histogram(dist, bins, range)[0]

followed by `res += r` (:180-190).  MDAnalysis and dask are not installable
here, so the worker body below is the restatement: `distance_array` from the C
oracle on ONE thread (MDAnalysis' serial backend, which pmda uses), then the
`d <= max_cutoff` mask of `_bruteforce_capped`, then the real `np.histogram`.
Workers are spawned interpreters, like dask's.  This is the CPU figure
BASELINE.json's configs[0] names ("n_jobs=4 on CPU").
"""
import multiprocessing as mp
import os
import time

import numpy as np


def _worker_block(args):
    pos, dims, i1, i2, nbins, rng, excl = args
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import c_oracle
    t0 = time.perf_counter()
    res = None
    box = None if dims is None else np.asarray(dims, dtype=np.float32)
    for f in range(len(pos)):
        g1, g2 = pos[f][i1], pos[f][i2]
        d = c_oracle.distance_array(g1, g2, box)        # float64 [n1, n2]
        mask = np.where(d <= rng[1])                    # _bruteforce_capped
        dist = d[mask]
        if excl is not None:
            keep = (mask[0] // excl[0]) != (mask[1] // excl[1])
            dist = dist[keep]
        count = np.histogram(dist, bins=nbins, range=rng)[0]
        res = count if res is None else res + count     # _reduce
    return res, time.perf_counter() - t0


def _warm(_):
    import numpy   # noqa: F401
    return os.getpid()


def interrdf_process_per_block(pos, dims, i1, i2, nbins, rng, n_jobs,
                               n_blocks=None, excl=None):
    """-> dict(count, seconds, seconds_with_startup, n_jobs, n_blocks).
    `seconds` is the wall time of the block map on warm workers (dask keeps
    its workers alive between runs); `seconds_with_startup` includes spawning
    them, as a first run() would see it."""
    from pmda_b200.util import make_balanced_slices
    n_blocks = n_blocks or n_jobs
    n_frames = len(pos)
    slices = make_balanced_slices(n_frames, n_blocks, start=0, stop=n_frames,
                                  step=1)
    blocks = [(np.ascontiguousarray(pos[sl]), dims, np.asarray(i1),
               np.asarray(i2), nbins, tuple(rng), excl) for sl in slices]
    ctx = mp.get_context("spawn")
    t_start = time.perf_counter()
    with ctx.Pool(processes=n_jobs) as pool:
        pool.map(_warm, range(4 * n_jobs))
        t0 = time.perf_counter()
        results = pool.map(_worker_block, blocks, chunksize=1)
        t1 = time.perf_counter()
    count = None
    for c, _ in results:
        if c is not None:
            count = c if count is None else count + c
    return {"count": count, "seconds": t1 - t0,
            "seconds_with_startup": t1 - t_start, "n_jobs": n_jobs,
            "n_blocks": n_blocks,
            "worker_seconds": [float(s) for _, s in results]}
