"""Per-GPU frame-block streams: what replaces pmda's dask task graph
(pmda/parallel.py:376-392) for the RDF analyses.

For every block of frames that ``make_balanced_slices`` produced, the engine

  1. cuts the block into batches of frames,
  2. stages each batch host -> pinned -> device on a copy stream (or takes it
     from the device-resident frame cache),
  3. launches the C-ABI block kernel (``rdfk_rdf_hist`` / ``rdfk_rdf_sites``)
     on the compute stream, accumulating into that block's device histogram,
  4. at the end reads all block histograms back with one D2H copy.

Frames are sharded over the local GPUs (``n_jobs``) and, under
``torch.distributed``, over ranks; partial int64 histograms are combined with a
single all-reduce.  Integer sums are order independent, so counts are
bit-identical for any GPU count; the float64 per-frame volumes are summed on
the host in frame order, like ``_reduce`` does in the reference.

There is no CPU fallback: without the CUDA library / a GPU this raises.
"""
import collections
import itertools
import os
import threading
import time
import warnings
import weakref
from collections import OrderedDict

import numpy as np
import torch

from . import _cabi, mdamath
from .util import make_balanced_slices


#: host-side trace points (tools/probe_host_time.py): a list collects
#: (label, perf_counter) pairs while it is installed here; None = off
_trace = None


def _tp(label):
    if _trace is not None:
        _trace.append((label, time.perf_counter()))


class Config(object):
    #: raw frame bytes staged per batch (bigger = fewer launches)
    batch_bytes = int(os.environ.get("PMDA_B200_BATCH_BYTES", 128 << 20))
    #: never more frames than this per kernel call
    max_batch_frames = int(os.environ.get("PMDA_B200_MAX_BATCH_FRAMES", 2048))
    #: device-resident frame cache budget per GPU (0 disables the cache)
    cache_bytes = int(os.environ.get("PMDA_B200_CACHE_BYTES", 48 << 30))
    #: compute streams the RDF batches alternate between (1 or 2)
    compute_streams = int(os.environ.get("PMDA_B200_COMPUTE_STREAMS", 2))
    #: uncached runs: first batch = 1/head_divisor of the frames (0: off)
    head_divisor = int(os.environ.get("PMDA_B200_HEAD_DIVISOR", 16))
    #: ... but only when the run uploads at least this many bytes: for a small
    #: run a second launch costs more than the upload it would hide
    head_min_bytes = int(os.environ.get("PMDA_B200_HEAD_MIN_BYTES", 64 << 20))
    #: host threads gathering the atoms an analysis references into the
    #: pinned ring (InterRDF_s: only those atoms cross PCIe)
    gather_threads = int(os.environ.get(
        "PMDA_B200_GATHER_THREADS",
        max(1, min(16, (os.cpu_count() or 1) //
                   max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))))
    #: page-lock in-memory trajectories so H2D runs straight from them
    pin_host = os.environ.get("PMDA_B200_PIN_HOST", "1") != "0"
    #: host threads copying frame blocks into the pinned ring
    copy_threads = int(os.environ.get("PMDA_B200_COPY_THREADS",
                                      min(8, os.cpu_count() or 1)))
    #: host threads hashing a cached frame block's host data (every byte; the
    #: hash runs while the GPU already works on the cached copy)
    hash_threads = int(os.environ.get(
        "PMDA_B200_HASH_THREADS",
        max(1, min(16, (os.cpu_count() or 1) //
                   max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))))


# --------------------------------------------------------------------------
# backends.  The product backend is CUDA; tests may install a stand-in to
# exercise the sharding logic on a CPU-only machine.
# --------------------------------------------------------------------------
class CudaBackend(object):
    name = "cuda"

    def devices(self, n_jobs):
        if not torch.cuda.is_available():
            raise _cabi.RdfkError(
                "pmda_b200 needs a CUDA device (sm_100a): there is no CPU "
                "fallback for InterRDF / InterRDF_s")
        _cabi.lib()
        if _dist_world() > 1:
            return [torch.device("cuda", torch.cuda.current_device())]
        n = max(1, min(int(n_jobs), torch.cuda.device_count()))
        first = torch.cuda.current_device()
        return [torch.device("cuda", (first + k) % torch.cuda.device_count())
                for k in range(n)]

    rdf_workspace_bytes = staticmethod(_cabi.rdf_workspace_bytes)
    rdf_hist = staticmethod(_cabi.rdf_hist)
    rdf_sites = staticmethod(_cabi.rdf_sites)
    contacts = staticmethod(_cabi.contacts)
    unpack_planes = staticmethod(_cabi.unpack_planes)
    hbonds_workspace_bytes = staticmethod(_cabi.hbonds_workspace_bytes)
    hbonds = staticmethod(_cabi.hbonds)
    host_hash_rows = staticmethod(_cabi.host_hash_rows)
    host_gather_rows = staticmethod(_cabi.host_gather_rows)
    host_sites_finish = staticmethod(_cabi.host_sites_finish)
    rdf_stats_async = staticmethod(_cabi.rdf_stats_async)
    zero_async = staticmethod(_cabi.zero_async)


_backend = CudaBackend()


def _set_test_backend(backend):
    """Test hook (tests/ only): swap the kernel backend; ``None`` restores CUDA."""
    global _backend
    _backend = CudaBackend() if backend is None else backend


def _dist_world():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and \
        dist.is_initialized() else 1


def _dist_rank():
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() \
        else 0


# --------------------------------------------------------------------------
# per-device state: streams, staging buffers, workspace, frame cache
# --------------------------------------------------------------------------
class _DeviceState(object):
    def __init__(self, device):
        self.device = device
        self.is_cuda = device.type == "cuda"
        self.lock = threading.Lock()
        self.workspaces = [None, None]
        #: which of the two compute streams (and workspaces) is in use
        self.cur = 0
        self.pinned = [None, None]
        self.staged = [None, None]
        self.raw_pin = [None, None]      # raw file bytes (DCD blocks)
        self.raw_dev = [None, None]
        self.cache = OrderedDict()   # key -> _CacheEntry
        self.cache_used = 0
        #: wrapped-frames counter of each workspace at the last report
        self.wrapped_seen = {}
        #: small device constants (edges, index lists, box rows) by content
        self.consts = OrderedDict()
        #: page-locked landing buffer of a run's results
        self.result_pin = None
        #: the device side of it (int64, grows): [counters | results | extras]
        self.result_dev = None
        if self.is_cuda:
            with torch.cuda.device(device):
                self.copy_stream = torch.cuda.Stream(device)
                # two compute streams: consecutive batches of the RDF path
                # alternate, so the next batch's kernels fill the SMs the
                # persistent pair kernel's last work items leave idle
                self.compute_streams = [torch.cuda.Stream(device),
                                        torch.cuda.Stream(device)]
        else:
            self.copy_stream = None
            self.compute_streams = [None, None]

    @property
    def compute_stream(self):
        return self.compute_streams[self.cur]

    def result_vector(self, n):
        """zeroed int64 device vector of n elements for this run's results
        (persistent storage; zeroed on the first compute stream)"""
        if not self.is_cuda:
            return torch.zeros(n, dtype=torch.int64)
        buf = self.result_dev
        if buf is None or buf.numel() < n:
            if buf is not None:
                for cs in self.compute_streams:
                    buf.record_stream(cs)
            self.result_dev = None
            buf = self.result_dev = torch.empty(max(n, 4096),
                                                dtype=torch.int64,
                                                device=self.device)
            # allocated on torch's current stream, used on ours from now on
            torch.cuda.current_stream(self.device).synchronize()
        res = buf[:n]
        _backend.zero_async(res, self.compute_streams[0].cuda_stream)
        return res

    # staging buffers are flat and only ever grow: analyses of different
    # shapes taking turns on a device must not re-allocate (page-locking
    # memory costs milliseconds)
    def pinned_view(self, slot, n_frames, n_atoms):
        need = max(1, n_frames * n_atoms * 3)
        buf = self.pinned[slot]
        if buf is None or buf.numel() < need:
            self.pinned[slot] = None
            buf = self.pinned[slot] = torch.empty(
                max(need, 1 << 16), dtype=torch.float32, pin_memory=True)
        return buf[:n_frames * n_atoms * 3].view(n_frames, n_atoms, 3)

    def staged_view(self, slot, n_frames, n_atoms):
        need = max(1, n_frames * n_atoms * 3)
        buf = self.staged[slot]
        if buf is None or buf.numel() < need:
            if buf is not None and self.is_cuda:
                # still read by copies / kernels queued on other streams
                for cs in self.compute_streams + [self.copy_stream]:
                    buf.record_stream(cs)
            self.staged[slot] = None
            buf = self.staged[slot] = torch.empty(
                max(need, 1 << 16), dtype=torch.float32, device=self.device)
        return buf[:n_frames * n_atoms * 3].view(n_frames, n_atoms, 3)

    def sync_compute(self):
        for cs in self.compute_streams:
            if cs is not None:
                cs.synchronize()

    def get_workspace(self, nbytes):
        """the workspace of the current compute stream"""
        have = self.workspaces[self.cur]
        if have is None or have.numel() < nbytes + 256:
            self.workspaces[self.cur] = None
            self.workspaces[self.cur] = torch.empty(
                int(nbytes) + 256, dtype=torch.uint8, device=self.device)
            # librdfk recognises its own header by a magic word: make sure a
            # recycled allocation does not carry an old one
            # (on the stream the kernels of this workspace run on)
            if self.is_cuda:
                with torch.cuda.stream(self.compute_stream):
                    self.workspaces[self.cur][:512].zero_()
            else:
                self.workspaces[self.cur][:512].zero_()
            self.wrapped_seen.pop(self.cur, None)
        ws = self.workspaces[self.cur]
        off = (-ws.data_ptr()) % 256
        return ws[off:off + nbytes]


_states = {}
_states_lock = threading.Lock()


def _state(device):
    key = (device.type, device.index)
    with _states_lock:
        st = _states.get(key)
        if st is None:
            st = _states[key] = _DeviceState(device)
        return st


def _evict(st, key=None):
    """Remove one cached frame block (`key`, or the least recently used one).
    The block was allocated on torch's current stream but is read by kernels
    on the compute streams, and the host runs batches ahead of the device:
    tell the caching allocator not to hand the memory out again before the
    work queued on those streams has finished."""
    if key is None:
        key, entry = st.cache.popitem(last=False)
    else:
        entry = st.cache.pop(key)
    if st.is_cuda and entry.xyz.is_cuda:
        for cs in st.compute_streams:
            entry.xyz.record_stream(cs)
    st.cache_used -= entry.nbytes


def clear_cache():
    """Drop every device-resident frame block (and the staging buffers)."""
    with _states_lock:
        for st in _states.values():
            while st.cache:
                _evict(st)
            st.cache_used = 0


# A cached frame block is identified by a token that is unique to the live
# trajectory object (never ``id()``: ids are reused after garbage collection)
# and the frame range.  It is only ever *served* if a 64-bit hash over EVERY
# byte of the host data it was uploaded from still matches (the reference
# re-reads each frame on each run, pmda/parallel.py:434: an in-place edit of
# an in-memory trajectory, or a rewritten file, must show up in the next
# run()).  Readers whose frames cannot be re-read cheaply (no ``frame_block``
# / ``raw_block``) are never cached.
_token_counter = itertools.count(1)


def _traj_token(traj):
    tok = getattr(traj, "_pmda_b200_token", None)
    if tok is None:
        tok = next(_token_counter)
        try:
            traj._pmda_b200_token = tok
            weakref.finalize(traj, _drop_token, tok)
        except (AttributeError, TypeError):
            return None          # cannot tag this reader: never cache it
    return tok


# Finalizers can run at any allocation (garbage collection), i.e. also while
# this very thread holds _states_lock or a device lock: they must not take
# locks.  They only record the dead token (deque.append is atomic); the cache
# entries are purged the next time the device state is used.
_dead_tokens = collections.deque()


def _drop_token(tok):
    _dead_tokens.append(tok)


def _purge_dead(st):
    """drop the cached frame blocks of trajectories that no longer exist
    (caller holds ``st.lock``)"""
    if not _dead_tokens:
        return
    dead = set(_dead_tokens)
    for key in [k for k in st.cache if k[0] in dead]:
        _evict(st, key)


class _CacheEntry(object):
    """one frame block resident on a device; ``hash`` = content hash of the
    host data it was uploaded from (an int, or the Future computing it)"""
    __slots__ = ("xyz", "nbytes", "hash")

    def __init__(self, xyz, nbytes, hash_):
        self.xyz, self.nbytes, self.hash = xyz, nbytes, hash_

    def hash_value(self):
        h = self.hash
        if not isinstance(h, int):
            h = self.hash = int(h.result())
        return h


class _StaleFrames(Exception):
    """a cached frame block did not match the trajectory any more; the stale
    blocks have been evicted and the part must run again"""


_hash_pool = None


def _block_hash(arr):
    """content hash of every byte of a block of frames (rows = frames)"""
    return _backend.host_hash_rows(arr, 0x706d6461, Config.hash_threads)


def _hash_async(arr):
    """-> Future of :func:`_block_hash`; the native hash releases the GIL, so
    it runs while this thread keeps launching kernels"""
    global _hash_pool
    if _hash_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _hash_pool = ThreadPoolExecutor(max_workers=2,
                                        thread_name_prefix="pmda-b200-hash")
    return _hash_pool.submit(_block_hash, arr)


_pinned_arrays = {}


def _maybe_pin(arr):
    """cudaHostRegister a trajectory array once so copies from it are async."""
    if not Config.pin_host or not torch.cuda.is_available():
        return
    base = arr
    while isinstance(getattr(base, "base", None), np.ndarray):
        base = base.base
    key = base.__array_interface__["data"][0]
    if key in _pinned_arrays or base.nbytes < (1 << 20):
        return
    if not base.flags.writeable:
        # read-only file mappings (NpyTrajectory) cannot be page-locked in
        # place; their blocks go through the pinned staging ring
        _pinned_arrays[key] = False
        return
    try:
        rc = torch.cuda.cudart().cudaHostRegister(key, base.nbytes, 0)
        ok = int(rc) == 0
    except Exception:       # registration is an optimisation only
        ok = False
    _pinned_arrays[key] = ok
    if ok:
        weakref.finalize(base, _unpin, key)


def _unpin(key):
    _pinned_arrays.pop(key, None)
    try:
        torch.cuda.cudart().cudaHostUnregister(key)
    except Exception:
        pass


# --------------------------------------------------------------------------
# trajectory access
# --------------------------------------------------------------------------
_copy_pool = None


def _copy_frames(dst, src):
    """``dst[...] = src`` frame-parallel: for a file-backed trajectory this
    copy *is* the file read (page cache -> pinned ring), and numpy releases
    the GIL while it runs, so a few threads multiply the ingest bandwidth."""
    global _copy_pool
    n = len(src)
    nthreads = min(Config.copy_threads, n)
    if nthreads <= 1 or src.nbytes < (8 << 20):
        np.copyto(dst, src)
        return
    if _copy_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _copy_pool = ThreadPoolExecutor(max_workers=Config.copy_threads)
    cuts = np.linspace(0, n, nthreads + 1).astype(int)
    futs = [_copy_pool.submit(np.copyto, dst[a:b], src[a:b])
            for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    for fu in futs:
        fu.result()


_reader_locks = weakref.WeakKeyDictionary()
_reader_locks_guard = threading.Lock()
_reader_lock_fallback = threading.Lock()


def _reader_lock(traj):
    """one lock per live reader object (one shared lock for readers that
    cannot be weakly referenced)"""
    with _reader_locks_guard:
        try:
            lk = _reader_locks.get(traj)
            if lk is None:
                lk = _reader_locks[traj] = threading.Lock()
            return lk
        except TypeError:
            return _reader_lock_fallback


def _read_frames(traj, frames):
    """-> (positions float32 [B,N,3], dimensions float32 [B,6] | None)"""
    fb = getattr(traj, "frame_block", None)
    if fb is not None:
        return fb(frames)
    pos, dims = [], []
    # Any MDAnalysis-like reader: ``traj[i]`` moves the reader's one file
    # cursor and refills its one shared ``ts``.  With n_jobs > 1 every device
    # thread reads from the same reader object, so the seek and the copy out
    # of ``ts`` are one critical section (the reference avoids the problem by
    # re-opening the trajectory in every worker, pmda/parallel.py:421).
    with _reader_lock(traj):
        for i in frames:
            ts = traj[i]
            pos.append(np.array(ts.positions, dtype=np.float32, copy=True))
            d = getattr(ts, "dimensions", None)
            dims.append(None if d is None else np.array(d, dtype=np.float32))
    pos = np.stack(pos) if pos else np.zeros((0, 0, 3), np.float32)
    if any(d is None for d in dims):
        return pos, None
    return pos, np.stack(dims)


def _dev_const(st, arr, dtype):
    """Device copy of a small host array, looked up by CONTENT: run() is
    called again and again with the same edges / AtomGroup indices / box rows
    (often on a fresh analysis object), and every pageable upload is a
    synchronous ~15 us round trip."""
    a = np.ascontiguousarray(arr, dtype=dtype)
    if a.nbytes > (64 << 20):
        return torch.as_tensor(a, device=st.device)
    key = (a.dtype.str, a.shape,
           _backend.host_hash_rows(a.reshape(1, -1).view(np.uint8), 1, 1))
    hit = st.consts.get(key)
    if hit is not None:
        st.consts.move_to_end(key)
        return hit
    t = torch.as_tensor(a if a.flags.writeable else a.copy(),
                        device=st.device)
    if st.is_cuda:
        # uploaded on torch's current stream; every later user runs on a
        # compute stream
        torch.cuda.current_stream(st.device).synchronize()
    st.consts[key] = t
    while len(st.consts) > 256:
        st.consts.popitem(last=False)
    return t


_box_cache = OrderedDict()


def _box_row(row):
    """(kind, box9, volume) of one `dimensions` row, remembered by content"""
    key = row.tobytes()
    hit = _box_cache.get(key)
    if hit is None:
        kind, b = mdamath.classify_box(row)
        vol = float(mdamath.box_volume(row)) if kind != _cabi.BOX_NONE else 0.0
        hit = (kind, np.asarray(b, dtype=np.float32), vol)
        _box_cache[key] = hit
        while len(_box_cache) > 1024:
            _box_cache.popitem(last=False)
    return hit


_box_rows_cache = OrderedDict()


def _box_rows(dims, n):
    """per-frame (kind, box9, volume) with one host evaluation per distinct
    box; whole batches are remembered by content (NVT runs ask for the same
    rows again and again)"""
    if dims is None:
        key = n
    else:
        dims = np.ascontiguousarray(dims)
        key = (n, dims.dtype.str, dims.tobytes()) if dims.nbytes <= 65536 \
            else None
    if key is not None:
        hit = _box_rows_cache.get(key)
        if hit is not None:
            return hit
    out = _box_rows_uncached(dims, n)
    if key is not None:
        for a in out:
            a.setflags(write=False)
        _box_rows_cache[key] = out
        while len(_box_rows_cache) > 64:
            _box_rows_cache.popitem(last=False)
    return out


def _box_rows_uncached(dims, n):
    kinds = np.zeros(n, dtype=np.int32)
    box9 = np.zeros((n, 9), dtype=np.float32)
    vols = np.zeros(n, dtype=np.float64)
    if dims is None:
        return kinds, box9, vols
    d = np.ascontiguousarray(dims)
    if len(d) and (d == d[0]).all():        # NVT: one box for the whole batch
        uniq, inv = d[:1], np.zeros(len(d), dtype=np.intp)
    else:
        uniq, inv = np.unique(d, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
    for u, row in enumerate(uniq):
        kind, b, vol = _box_row(np.ascontiguousarray(row, dtype=np.float32))
        sel = inv == u
        kinds[sel] = kind
        box9[sel] = b
        vols[sel] = vol
    return kinds, box9, vols


def _runs(kinds):
    """consecutive runs of equal box kind -> [(kind, lo, hi)]"""
    kinds = np.asarray(kinds)
    n = len(kinds)
    if n == 0:
        return []
    k0 = kinds[0]
    if k0 == kinds[n - 1] and (kinds == k0).all():   # the usual case: one run
        return [(int(k0), 0, n)]
    cuts = np.flatnonzero(kinds[1:] != kinds[:-1]) + 1
    lo = np.concatenate(([0], cuts))
    hi = np.concatenate((cuts, [len(kinds)]))
    return [(int(kinds[a]), int(a), int(b)) for a, b in zip(lo, hi)]


def _frame_shards(blocks, n_parts):
    """Shard the concatenated frame list of all blocks into ``n_parts``
    contiguous, balanced pieces (same rule as the block partition).  Returns
    for every part a list of (block_id, position_in_run, frames: range)."""
    sizes = [len(b) for b in blocks]
    total = sum(sizes)
    parts = [[] for _ in range(n_parts)]
    if total == 0:
        return parts
    n_used = min(n_parts, total)
    cuts = make_balanced_slices(total, n_used)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    for p, sl in enumerate(cuts):
        lo = sl.start
        hi = total if sl.stop is None else sl.stop
        for b, blk in enumerate(blocks):
            a, z = max(lo, starts[b]), min(hi, starts[b + 1])
            if a < z:
                sub = blk[int(a - starts[b]):int(z - starts[b])]
                parts[p].append((b, int(a), sub))
    return parts


# --------------------------------------------------------------------------
# the staging pipeline of one device
# --------------------------------------------------------------------------
class _Pipeline(object):
    """Feeds frame batches of one trajectory to one device, double buffered."""

    def __init__(self, st, traj, natoms, gather=None):
        """`gather`: sorted int32 atom indices -- only these atoms go to the
        device (as xyz [B, len(gather), 3], in this order)"""
        self.st, self.traj = st, traj
        self.natoms_traj = natoms
        self.gather = gather
        if gather is not None:
            natoms = len(gather)
        self.natoms = natoms
        self.frame_bytes = max(1, natoms) * 12
        nb = max(1, Config.batch_bytes // max(1, self.frame_bytes))
        self.batch_frames = int(min(nb, Config.max_batch_frames))
        self.dslot = 0
        self.pslot = 0
        self.h2d_done = [None, None]
        self.compute_done = [None, None]
        _purge_dead(st)
        #: in-memory / memory-mapped trajectories hand out views: re-reading
        #: them is free, so a cached block is validated against a hash of the
        #: current data on every hit.  Other readers are never cached.
        self.cheap_reads = getattr(traj, "frame_block", None) is not None
        #: file formats whose frames can go to the device as they lie on disk
        self.raw = getattr(traj, "raw_block", None) \
            if (st.is_cuda and gather is None) else None
        # Gathered blocks are never cached: validating a cached copy means
        # reading every referenced atom of the host data again, which is the
        # gather itself -- re-gathering costs the same host pass and the
        # (3x smaller) upload hides behind it.
        self.traj_key = _traj_token(traj) \
            if (gather is None and
                (self.cheap_reads or self.raw is not None)) else None
        #: cache hits of this run whose host data is still being hashed:
        #: (key, entry, Future of the current data's hash)
        self.pending = []
        self.rslot = 0
        self.raw_done = [None, None]
        self.h2d_raw_done = [None, None]

    def batches(self, frames):
        """equal-sized batches of at most ``batch_frames`` frames (a short
        last batch would leave the persistent kernel's tail under-filled),
        after a short head batch when the frames stream from the host"""
        n = len(frames)
        if n == 0:
            return
        head = 0
        streaming = self.st.is_cuda and (self.traj_key is None
                                         or Config.cache_bytes <= 0)
        if streaming and Config.head_divisor > 0 \
                and n >= 4 * Config.head_divisor \
                and n * self.frame_bytes >= Config.head_min_bytes:
            # frames come from the host every time: a short first batch, whose
            # upload is the only one no kernel hides (its own kernel tail is
            # covered by the next batch on the other compute stream)
            head = min(n // Config.head_divisor, self.batch_frames)
            yield frames[:head]
        rest = n - head
        nb = -(-rest // self.batch_frames)
        if self.gather is not None and streaming and rest >= 64:
            # the host gather of batch k+1 runs beside the upload and the
            # kernels of batch k: a few batches even when one would hold all
            nb = max(nb, 4)
        size = -(-rest // nb)
        for lo in range(head, n, size):
            yield frames[lo:lo + size]

    def fetch(self, frames):
        """-> (xyz_dev [B,N,3], dims_host, io_seconds, device slot or None)"""
        st = self.st
        t0 = time.time()
        use_cache = self.traj_key is not None and Config.cache_bytes > 0
        pos = dims = None
        key = None
        if self.gather is not None:
            return self._fetch_gather(frames, t0)
        rb = self.raw(frames) if self.raw is not None else None
        if rb is not None:
            return self._fetch_raw(frames, rb, use_cache, t0)
        hfut = None
        if use_cache:
            pos, dims = _read_frames(self.traj, frames)
            key = (self.traj_key, frames.start, frames.stop, frames.step,
                   self.natoms)
            # every byte of the host data is hashed -- on a helper thread,
            # while this one already launches kernels on the cached copy;
            # verify() compares before any result of the run is used
            hfut = _hash_async(pos)
            hit = st.cache.get(key)
            if hit is not None:
                st.cache.move_to_end(key)
                self.pending.append((key, hit, hfut))
                return hit.xyz, dims, time.time() - t0, None
        if pos is None:
            pos, dims = _read_frames(self.traj, frames)
        B = len(frames)
        if pos.shape[1:] != (self.natoms, 3):
            raise ValueError("trajectory frame has %r atoms, expected %d"
                             % (pos.shape[1:], self.natoms))
        nbytes = B * self.frame_bytes
        cacheable = use_cache and nbytes <= Config.cache_bytes
        if cacheable:
            while st.cache and st.cache_used + nbytes > Config.cache_bytes:
                _evict(st)                                        # LRU
        if not st.is_cuda:
            # stand-in backends of the tests: same cache logic, host tensors
            host = np.ascontiguousarray(pos)
            if cacheable or not host.flags.writeable:   # own copy / mapping
                host = np.array(host)
            xyz_dev = torch.from_numpy(host)
            if cacheable:
                st.cache[key] = _CacheEntry(xyz_dev, nbytes, hfut)
                st.cache_used += nbytes
            return xyz_dev, dims, time.time() - t0, None
        slot = None
        with torch.cuda.device(st.device):
            # destination: a cache entry of its own, or a rotating buffer
            if cacheable:
                xyz_dev = torch.empty((B, self.natoms, 3), dtype=torch.float32,
                                      device=st.device)
            else:
                slot = self.dslot
                self.dslot ^= 1
                if self.compute_done[slot] is not None:
                    st.copy_stream.wait_event(self.compute_done[slot])
                xyz_dev = st.staged_view(slot, B, self.natoms)
            # source: the trajectory array itself when it is page-locked,
            # otherwise a rotating pinned staging buffer
            src = None
            if pos.flags.c_contiguous and pos.flags.writeable:
                _maybe_pin(pos)
                cand = torch.from_numpy(pos)
                if cand.is_pinned():
                    src = cand
            if src is None:
                ps = self.pslot
                self.pslot ^= 1
                if self.h2d_done[ps] is not None:
                    self.h2d_done[ps].synchronize()
                src = st.pinned_view(ps, B, self.natoms)
                _copy_frames(src.numpy(), pos)
            else:
                ps = None
            with torch.cuda.stream(st.copy_stream):
                xyz_dev.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(st.copy_stream)
            if ps is not None:
                self.h2d_done[ps] = ev
            st.compute_stream.wait_event(ev)
        if cacheable:
            st.cache[key] = _CacheEntry(xyz_dev, nbytes, hfut)
            st.cache_used += nbytes
        return xyz_dev, dims, time.time() - t0, slot

    def _fetch_gather(self, frames, t0):
        """only the atoms of `self.gather`: host gather (threaded, native)
        straight into the pinned ring, upload from there"""
        st = self.st
        pos, dims = _read_frames(self.traj, frames)
        B = len(frames)
        if pos.shape[1:] != (self.natoms_traj, 3):
            raise ValueError("trajectory frame has %r atoms, expected %d"
                             % (pos.shape[1:], self.natoms_traj))
        if pos.dtype != np.float32 or (B and not pos[0].flags.c_contiguous):
            pos = np.ascontiguousarray(pos, dtype=np.float32)
        ng = self.natoms
        if not st.is_cuda:
            host = np.empty((B, ng, 3), dtype=np.float32)
            _backend.host_gather_rows(pos, self.gather, host,
                                      Config.gather_threads)
            return torch.from_numpy(host), dims, time.time() - t0, None
        with torch.cuda.device(st.device):
            ps = self.pslot
            self.pslot ^= 1
            if self.h2d_done[ps] is not None:
                self.h2d_done[ps].synchronize()         # pinned slot is free
            pin = st.pinned_view(ps, B, ng)
            _backend.host_gather_rows(pos, self.gather, pin.numpy(),
                                      Config.gather_threads)
            slot = self.dslot
            self.dslot ^= 1
            if self.compute_done[slot] is not None:
                st.copy_stream.wait_event(self.compute_done[slot])
            xyz_dev = st.staged_view(slot, B, ng)
            with torch.cuda.stream(st.copy_stream):
                xyz_dev.copy_(pin, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(st.copy_stream)
            self.h2d_done[ps] = ev
            st.compute_stream.wait_event(ev)
        return xyz_dev, dims, time.time() - t0, slot

    def verify(self):
        """Compare the hashes of this run's cache hits with the hashes of the
        data the cached blocks were uploaded from.  Stale blocks are evicted
        and :class:`_StaleFrames` tells the caller to run the part again (the
        kernels already ran on the stale copy; their results are dropped)."""
        pending, self.pending = self.pending, []
        stale = []
        for key, entry, fut in pending:
            if int(fut.result()) != entry.hash_value():
                stale.append((key, entry))
        for key, entry in stale:
            if self.st.cache.get(key) is entry:
                _evict(self.st, key)
        if stale:
            raise _StaleFrames()

    def _fetch_raw(self, frames, rb, use_cache, t0):
        """frames as they lie in the file -> pinned ring -> device ->
        rdfk_unpack_planes -> xyz [B,N,3]"""
        st = self.st
        raw, fstride, offs, dims = rb
        B = raw.shape[0]
        nbytes = B * self.frame_bytes
        key = None
        hfut = None
        if use_cache:
            key = (self.traj_key, frames.start, frames.stop, frames.step,
                   self.natoms)
            hfut = _hash_async(raw)          # every byte of the file block
            hit = st.cache.get(key)
            if hit is not None:
                st.cache.move_to_end(key)
                self.pending.append((key, hit, hfut))
                return hit.xyz, dims, time.time() - t0, None
        cap = self.batch_frames * fstride
        nraw = B * fstride
        rs = self.rslot
        self.rslot ^= 1
        with torch.cuda.device(st.device):
            for bufs, kw in ((st.raw_pin, {"pin_memory": True}),
                             (st.raw_dev, {"device": st.device})):
                if bufs[rs] is None or bufs[rs].numel() < cap:
                    bufs[rs] = None
                    bufs[rs] = torch.empty(cap, dtype=torch.uint8, **kw)
            if self.h2d_raw_done[rs] is not None:
                self.h2d_raw_done[rs].synchronize()     # pinned slot is free
            _copy_frames(st.raw_pin[rs][:nraw].numpy().reshape(B, fstride), raw)
            if self.raw_done[rs] is not None:           # device slot is free
                st.copy_stream.wait_event(self.raw_done[rs])
            with torch.cuda.stream(st.copy_stream):
                st.raw_dev[rs][:nraw].copy_(st.raw_pin[rs][:nraw],
                                            non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(st.copy_stream)
            self.h2d_raw_done[rs] = ev
            st.compute_stream.wait_event(ev)
            cacheable = use_cache and nbytes <= Config.cache_bytes
            slot = None
            if cacheable:
                while st.cache and st.cache_used + nbytes > Config.cache_bytes:
                    _evict(st)
                xyz_dev = torch.empty((B, self.natoms, 3), dtype=torch.float32,
                                      device=st.device)
            else:
                slot = self.dslot
                self.dslot ^= 1
                xyz_dev = st.staged_view(slot, B, self.natoms)
            # the transpose runs on the compute stream, in order with the
            # kernels that read (and last read) xyz_dev
            _backend.unpack_planes(st.raw_dev[rs], fstride, offs[0], offs[1],
                                   offs[2], B, self.natoms, xyz_dev,
                                   st.compute_stream.cuda_stream)
            evu = torch.cuda.Event()
            evu.record(st.compute_stream)
            self.raw_done[rs] = evu
        if cacheable:
            st.cache[key] = _CacheEntry(xyz_dev, nbytes, hfut)
            st.cache_used += nbytes
        return xyz_dev, dims, time.time() - t0, slot

    def mark_compute_done(self, slot):
        if slot is None or not self.st.is_cuda:
            return
        ev = torch.cuda.Event()
        ev.record(self.st.compute_stream)
        self.compute_done[slot] = ev


class _Timer(object):
    """CUDA-event timing of one batch on the compute stream."""

    def __init__(self, st):
        self.st = st
        if st.is_cuda:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
        else:
            self.t0 = self.t1 = 0.0

    def start(self):
        if self.st.is_cuda:
            self.e0.record(self.st.compute_stream)
        else:
            self.t0 = time.time()

    def stop(self):
        if self.st.is_cuda:
            self.e1.record(self.st.compute_stream)
        else:
            self.t1 = time.time()

    def seconds(self, after=None):
        """elapsed seconds; with `after` (the previous batch's timer) only the
        part that follows that batch's end"""
        if self.st.is_cuda:
            self.e1.synchronize()
            ms = self.e0.elapsed_time(self.e1)
            if after is not None:
                ms = max(0.0, min(ms, after.e1.elapsed_time(self.e1)))
            return ms * 1e-3
        return self.t1 - self.t0


# --------------------------------------------------------------------------
# public entry points
# --------------------------------------------------------------------------
def _to_dev_i32(idx, device):
    return torch.as_tensor(np.asarray(idx, dtype=np.int32), device=device)


def _as_u64_ptr_tensor(t):
    return t  # int64 storage, the kernels treat it as uint64


def _result_to_host(st, res, stream):
    """one D2H copy of a run's device results into the device's page-locked
    landing buffer, behind everything queued on `stream` -> int64 numpy view
    (valid until the next run on this device)"""
    n = res.numel()
    if not st.is_cuda:
        return res.numpy()
    pin = st.result_pin
    if pin is None or pin.numel() < n:
        st.result_pin = None
        pin = st.result_pin = torch.empty(max(n, 4096), dtype=torch.int64,
                                          pin_memory=True)
    with torch.cuda.stream(stream):
        pin[:n].copy_(res, non_blocking=True)
    stream.synchronize()
    return pin[:n].numpy()


def _run_rdf_part(device, traj, natoms, segs, n_blocks, n_frames, idx1, idx2,
                  same_group, nbins, rng, edges, excl, n_extra=0):
    """All frames of one local device: stage the batches and queue the
    kernels.  Nothing here waits for the GPU.  -> dict of partial results;
    ``res`` is a device int64 vector ``[16 | n_blocks * nbins | n_extra]``:
    the counters of the two workspaces (rdfk.h, rdfk_rdf_stats), the block
    histograms, and room for what the caller wants to send along in the one
    all-reduce of the run."""
    st = _state(device)
    be = _backend
    out = {
        "res": None, "timers": [], "pipe": None, "st": st,
        "vol": np.zeros(n_frames, dtype=np.float64),
        "io": np.zeros(n_frames, dtype=np.float64),
        "compute": np.zeros(n_frames, dtype=np.float64),
        "first_launch": np.full(n_blocks, np.inf),
    }
    with st.lock:
        cuda = st.is_cuda
        ctx = torch.cuda.device(device) if (
            cuda and torch.cuda.current_device() != device.index) \
            else _nullctx()
        with ctx:
            n_streams = max(1, min(2, Config.compute_streams)) if cuda else 1
            st.cur = 0
            cs = st.compute_streams
            res = st.result_vector(16 + n_blocks * nbins + n_extra)
            hist = res[16:16 + n_blocks * nbins].view(n_blocks, nbins)
            _tp("part: result buffer")
            edges_d = _dev_const(st, edges, np.float64)
            i1 = _dev_const(st, idx1, np.int32)
            i2 = i1 if same_group else _dev_const(st, idx2, np.int32)
            n1, n2 = len(idx1), len(idx2)
            pipe = _Pipeline(st, traj, natoms)
            ws_bytes = be.rdf_workspace_bytes(pipe.batch_frames, n1, n2, nbins,
                                              same_group)
            timers = out["timers"]
            if cuda and n_streams > 1:
                cs[1].wait_stream(cs[0])          # the zero-fill of `res`
            _tp("part: constants, pipeline")
            n_batch = 0
            for (b, pos0, frames) in segs:
                done = 0
                for batch in pipe.batches(frames):
                    B = len(batch)
                    st.cur = n_batch % n_streams
                    n_batch += 1
                    ws = st.get_workspace(ws_bytes)
                    xyz, dims, t_io, slot = pipe.fetch(batch)
                    _tp("part: fetch")
                    kinds, box9, vols = _box_rows(dims, B)
                    lo_f = pos0 + done
                    out["vol"][lo_f:lo_f + B] = vols
                    out["io"][lo_f:lo_f + B] = t_io / B
                    box_d = _dev_const(st, box9, np.float32)
                    out["first_launch"][b] = min(out["first_launch"][b],
                                                 time.time())
                    _tp("part: box rows")
                    tm = _Timer(st)
                    tm.start()
                    stream = st.compute_stream.cuda_stream if cuda else 0
                    hist_b = hist[b]
                    for kind, lo, hi in _runs(kinds):
                        # (only the pointers travel: no upper bounds, and no
                        # tensor views at all for the usual one-kind batch)
                        be.rdf_hist(xyz if lo == 0 else xyz[lo:], hi - lo,
                                    natoms, i1, n1, i2, n2, same_group, kind,
                                    box_d if lo == 0 else box_d[lo:],
                                    rng[0], rng[1], nbins, edges_d, excl,
                                    hist_b, ws, stream)
                    tm.stop()
                    _tp("part: rdfk_rdf_hist (launches)")
                    pipe.mark_compute_done(slot)
                    timers.append((tm, lo_f, B))
                    done += B
            if cuda:
                # the counters librdfk keeps in the first 64 bytes of each
                # workspace travel with the histograms
                for k in range(min(n_streams, n_batch)):
                    ws_k = st.workspaces[k]
                    off = (-ws_k.data_ptr()) % 256
                    be.rdf_stats_async(ws_k[off:], res[8 * k:8 * k + 8],
                                       cs[k].cuda_stream)
                if n_streams > 1 and n_batch > 1:
                    cs[0].wait_stream(cs[1])      # cs[0] now ends the run
            st.cur = 0
            out["res"], out["pipe"] = res, pipe
            _tp("part: counters queued")
    return out


def _part_timers(out):
    """CUDA-event seconds of a part's batches -> out["compute"] (waits for
    the batches)"""
    prev = None
    for tm, lo_f, B in out["timers"]:
        # batches on alternating streams overlap: a batch is charged from
        # the end of the previous one (or its own start, if later)
        out["compute"][lo_f:lo_f + B] = tm.seconds(after=prev) / B
        prev = tm
    out["timers"] = []


def _check_stats(st, stats_host):
    """act on librdfk's workspace counters after a run (rdfk.h): internal
    violations are fatal; atoms outside the primary triclinic cell are
    reported once per growth of the counter.  `stats_host`: one int64[8]
    per workspace."""
    for k, s in enumerate(stats_host):
        if s[6]:
            raise _cabi.RdfkError(
                "librdfk internal check failed %d time(s), first code %d: the "
                "histogram of this run is not trustworthy" % (s[6], s[7]))
        seen = st.wrapped_seen.get(k, 0)
        if s[5] > seen:
            st.wrapped_seen[k] = int(s[5])
            warnings.warn(
                "%d frame(s) held atoms outside the primary triclinic unit "
                "cell (%d atoms so far); they were wrapped into it before the "
                "distance search like MDAnalysis' distance_array does, but "
                "the operation order of that wrap is the one part of the "
                "arithmetic that is not pinned against MDAnalysis (DESIGN.md "
                "section 6): wrap the trajectory beforehand to be on the "
                "pinned branch" % (s[5] - seen, s[4]), RuntimeWarning,
                stacklevel=4)


class _nullctx(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _combine(parts, device0, n_blocks, width, dtype):
    """sum partial [n_blocks, width] tensors of the local devices on device0"""
    total = None
    for p in parts:
        h = p["hist"].to(device0)
        total = h if total is None else total + h
    if total is None:
        total = torch.zeros((n_blocks, width), dtype=dtype, device=device0)
    return total


def _allreduce(t):
    import torch.distributed as dist
    if _dist_world() > 1:
        if t.is_cuda and dist.get_backend() == "gloo":
            c = t.cpu()
            dist.all_reduce(c)
            t.copy_(c)
        elif not t.is_cuda and dist.get_backend() == "nccl":
            c = t.cuda()
            dist.all_reduce(c)
            t.copy_(c.cpu())
        else:
            dist.all_reduce(t)
    return t


def _run_parts(fn_once, devices, parts):
    def fn(device, part):
        # a cached frame block turned out stale (the trajectory was edited
        # since it was uploaded): it has been evicted, run the part again
        for _ in range(4):
            try:
                return fn_once(device, part)
            except _StaleFrames:
                continue
        raise RuntimeError("the trajectory kept changing while run() was "
                           "reading it")

    if len(devices) == 1:
        return [fn(devices[0], parts[0])]
    results = [None] * len(devices)
    errors = []

    def work(k):
        try:
            results[k] = fn(devices[k], parts[k])
        except BaseException as exc:  # re-raised on the caller's thread
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(k,))
               for k in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def _shard_for_this_rank(blocks, devices):
    """frames -> ranks -> local devices; returns one segment list per device
    and the number of frames in the run."""
    world, rank = _dist_world(), _dist_rank()
    n_frames = sum(len(b) for b in blocks)
    if world == 1 and len(devices) == 1:      # everything is here
        segs, pos = [], 0
        for b, blk in enumerate(blocks):
            if len(blk):
                segs.append((b, pos, blk))
            pos += len(blk)
        return [segs], n_frames
    mine = _frame_shards(blocks, world)[rank]
    # second level: this rank's frames over its local devices
    sub_blocks = [f for (_, _, f) in mine]
    local = _frame_shards(sub_blocks, len(devices))
    parts = []
    for segs in local:
        part = []
        for (sb, pos_in_mine, frames) in segs:
            b, pos0, parent = mine[sb]
            off = sum(len(x) for x in sub_blocks[:sb])
            part.append((b, pos0 + (pos_in_mine - off), frames))
        parts.append(part)
    return parts, n_frames


def _finish(hist_np, vol, io, compute, first_launch, blocks, make_result):
    """per-block tuples in the layout pmda/parallel.py:442-444 returns"""
    res = []
    pos = 0
    for b, blk in enumerate(blocks):
        n = len(blk)
        # sequential, like `res += r` frame by frame (np.add.accumulate is a
        # strict left-to-right scan: the same additions in the same order)
        v = float(np.add.accumulate(vol[pos:pos + n])[-1]) if n else 0.0
        t_io = io[pos:pos + n].copy()
        t_c = compute[pos:pos + n].copy()
        wait_end = first_launch[b] if np.isfinite(first_launch[b]) \
            else time.time()
        res.append((make_result(b, hist_np[b], np.float64(v)), t_io, t_c, 0.0,
                    wait_end, np.sum(t_io), np.sum(t_c)))
        pos += n
    return res


def _local_parts(fn, devices, parts, fetch_now):
    """Queue the parts on the local devices, wait for them and make sure
    every cached frame block they used still matches the trajectory (stale
    blocks are evicted and everything runs again: the kernels had already
    consumed them).  With `fetch_now` (one device, one rank: nothing to
    exchange) the device results come to the host behind the last kernel, so
    the run has a single synchronisation.  -> (outs, host result or None)"""
    for _ in range(4):
        outs = _run_parts(fn, devices, parts)
        host = None
        if fetch_now:
            st = outs[0]["st"]
            with st.lock:
                host = _result_to_host(st, outs[0]["res"],
                                       st.compute_streams[0])
        else:
            for o in outs:
                if o["st"].is_cuda:
                    o["st"].compute_streams[0].synchronize()
        _tp("local: device finished")
        try:
            for o in outs:
                with o["st"].lock:
                    o["pipe"].verify()
        except _StaleFrames:
            continue
        for o in outs:
            _part_timers(o)
        return outs, host
    raise RuntimeError("the trajectory kept changing while run() was "
                       "reading it")


def _exchange(outs, host, n_main, n_frames):
    """The epilogue of a run: per-frame volumes / staging / kernel seconds of
    all local parts join the device results and cross the ranks in ONE
    all-reduce (float64 bit patterns ride along as int64: every frame is
    owned by exactly one rank, the others contribute zeros, and x + 0 is
    exact on the bits); one D2H copy.  -> (main int64 [n_main] of the whole
    job, extras float64 [3, n_frames], per-workspace counters of this rank)"""
    world = _dist_world()
    ext = np.stack([sum(o["vol"] for o in outs), sum(o["io"] for o in outs),
                    sum(o["compute"] for o in outs)])
    if host is not None:                  # one device, one rank: all here
        return host[16:16 + n_main], ext, [host[0:8], host[8:16]]
    stats = []
    if len(outs) == 1:
        o = outs[0]
        st, res = o["st"], o["res"]
        cs0 = st.compute_streams[0]
        with st.lock:
            with (torch.cuda.stream(cs0) if st.is_cuda else _nullctx()):
                if world > 1:
                    bits = torch.from_numpy(
                        np.ascontiguousarray(ext).reshape(-1).view(np.int64))
                    res[16 + n_main:].copy_(bits)
                    _allreduce(res[16:])
            full = _result_to_host(st, res, cs0)
        stats = [full[0:8], full[8:16]]
        main = full[16:16 + n_main]
        if world > 1:
            ext = full[16 + n_main:].view(np.float64).reshape(3, n_frames)
        return main, ext, stats
    # several local devices (n_jobs > 1): summed on the host
    total = np.zeros(n_main + 3 * n_frames, dtype=np.int64)
    for o in outs:
        with o["st"].lock:
            full = _result_to_host(o["st"], o["res"],
                                   o["st"].compute_streams[0]).copy()
        _check_stats(o["st"], [full[0:8], full[8:16]])
        total[:n_main] += full[16:16 + n_main]
    total[n_main:] = np.ascontiguousarray(ext).reshape(-1).view(np.int64)
    if world > 1:
        total = _allreduce(torch.from_numpy(total)).numpy()
    return (total[:n_main],
            total[n_main:].view(np.float64).reshape(3, n_frames), [])


def run_rdf_blocks(traj, natoms, idx1, idx2, same_group, blocks, nbins, rng,
                   edges, excl, n_jobs):
    """InterRDF over frame blocks.  -> list of per-block tuples whose first
    element is ``np.array([count int64[nbins], volume float64], dtype=object)``
    (the accumulated ``_single_frame`` results of pmda/rdf.py:137 + :180-190).
    """
    devices = _backend.devices(n_jobs)
    n_blocks = len(blocks)
    parts, n_frames = _shard_for_this_rank(blocks, devices)
    world = _dist_world()
    n_extra = 3 * n_frames if world > 1 else 0

    def fn(device, segs):
        return _run_rdf_part(device, traj, natoms, segs, n_blocks, n_frames,
                             idx1, idx2, same_group, nbins, rng, edges, excl,
                             n_extra)

    _tp("blocks: sharding")
    outs, host = _local_parts(fn, devices, parts,
                              fetch_now=(world == 1 and len(devices) == 1))
    _tp("blocks: wait for the device, verify, timers")
    main, extra, stats = _exchange(outs, host, n_blocks * nbins, n_frames)
    if stats:
        _check_stats(outs[0]["st"], stats)
    first = np.min(np.stack([o["first_launch"] for o in outs]), axis=0)
    hist_np = main.reshape(n_blocks, nbins)

    def make_result(b, count, vol):
        r = np.empty(2, dtype=object)
        r[0] = count.astype(np.int64, copy=True)
        r[1] = np.array(vol, dtype=np.float64)
        return r

    res = _finish(hist_np, extra[0], extra[1], extra[2], first, blocks,
                  make_result)
    _tp("blocks: results assembled")
    return res


def _site_gather(site, natoms):
    """The atoms the site pairs reference, sorted and unique, and the site
    index lists re-expressed as positions in that list -- or (None, idx_a,
    idx_b) when most of the frame is referenced anyway.  The reference
    gathers ``ag.positions`` first as well (pmda/rdf.py:296-301).  Kept in
    the site description: computed once per analysis object."""
    hit = site.get("_gather")
    if hit is not None and hit[0] == natoms:
        return hit[1]
    idx_a = np.asarray(site["idx_a"], dtype=np.int64)
    idx_b = np.asarray(site["idx_b"], dtype=np.int64)
    both = np.concatenate([idx_a, idx_b])
    if len(both) and (both.min() < 0 or both.max() >= natoms):
        raise IndexError("site atom index out of range")
    mark = np.zeros(natoms, dtype=bool)
    mark[both] = True
    used = np.flatnonzero(mark)
    if len(used) == 0 or 2 * len(used) > natoms:
        out = (None, idx_a, idx_b)
    else:
        where = np.empty(natoms, dtype=np.int32)
        where[used] = np.arange(len(used), dtype=np.int32)
        out = (used.astype(np.int32), where[idx_a], where[idx_b])
    site["_gather"] = (natoms, out)
    return out


def _run_sites_part(device, traj, natoms, segs, n_blocks, n_frames, site,
                    nbins, rng, edges):
    """All frames of one local device; like _run_rdf_part nothing here waits
    for the GPU.  ``res``: device int32 ``[n_blocks, total_cells * nbins]``."""
    st = _state(device)
    be = _backend
    total_cells = int(site["cell_off"][-1])
    out = {
        "res": None, "timers": [], "pipe": None, "st": st,
        "vol": np.zeros(n_frames, dtype=np.float64),
        "io": np.zeros(n_frames, dtype=np.float64),
        "compute": np.zeros(n_frames, dtype=np.float64),
        "first_launch": np.full(n_blocks, np.inf),
    }
    gather, idx_a, idx_b = _site_gather(site, natoms)
    with st.lock:
        cuda = st.is_cuda
        ctx = torch.cuda.device(device) if cuda else _nullctx()
        with ctx:
            st.cur = 0
            cs0 = st.compute_streams[0]
            with (torch.cuda.stream(cs0) if cuda else _nullctx()):
                count = torch.zeros((n_blocks, total_cells * nbins),
                                    dtype=torch.int32, device=device)
            edges_d = _dev_const(st, edges, np.float64)
            # the four int32 index lists travel (and are looked up) as one
            na, nb_, ns1 = len(idx_a), len(idx_b), len(site["off_a"])
            packed = _dev_const(st, np.concatenate(
                [idx_a, idx_b, site["off_a"], site["off_b"]]), np.int32)
            ia, ib = packed[:na], packed[na:na + nb_]
            oa = packed[na + nb_:na + nb_ + ns1]
            ob = packed[na + nb_ + ns1:]
            co = _dev_const(st, site["cell_off"], np.int64)
            pipe = _Pipeline(st, traj, natoms, gather=gather)
            timers = out["timers"]
            _tp("part: constants, pipeline")
            for (b, pos0, frames) in segs:
                done = 0
                for batch in pipe.batches(frames):
                    B = len(batch)
                    xyz, dims, t_io, slot = pipe.fetch(batch)
                    _tp("part: fetch")
                    kinds, box9, vols = _box_rows(dims, B)
                    lo_f = pos0 + done
                    out["vol"][lo_f:lo_f + B] = vols
                    out["io"][lo_f:lo_f + B] = t_io / B
                    box_d = _dev_const(st, box9, np.float32)
                    out["first_launch"][b] = min(out["first_launch"][b],
                                                 time.time())
                    _tp("part: box rows")
                    tm = _Timer(st)
                    tm.start()
                    stream = cs0.cuda_stream if cuda else 0
                    for kind, lo, hi in _runs(kinds):
                        be.rdf_sites(xyz[lo:hi], hi - lo, pipe.natoms, ia, ib,
                                     oa, ob, co, site["n_sites"], total_cells,
                                     kind, box_d[lo:hi], rng[0], rng[1], nbins,
                                     edges_d, count[b], stream)
                    tm.stop()
                    pipe.mark_compute_done(slot)
                    _tp("part: rdfk_rdf_sites (launch)")
                    timers.append((tm, lo_f, B))
                    done += B
            out["res"], out["pipe"] = count, pipe
    return out


def run_rdf_s_blocks(traj, natoms, site, blocks, nbins, rng, edges, n_jobs,
                     den_fn=None):
    """InterRDF_s over frame blocks.  ``site`` holds the CSR description of
    the site pairs (see rdf.InterRDF_s).  The first element of every block
    tuple is ``np.array([<object array of n float64 [n1,n2,nbins]>, volume])``
    like pmda/rdf.py:309 accumulated by :362-372.

    The float64 tensors are made from the int32 device counts by one threaded
    native pass (rdfk_host_sites_finish).  With ``den_fn`` (maps the list of
    per-block volumes to the rdf denominator ``[nbins]``, in the reference's
    own operation order) that pass also sums the blocks and divides:
    -> ``(per-block tuples, {"count": [...], "rdf": [...], "den": den})``.
    """
    devices = _backend.devices(n_jobs)
    n_blocks = len(blocks)
    total_cells = int(site["cell_off"][-1])
    n = total_cells * nbins
    parts, n_frames = _shard_for_this_rank(blocks, devices)
    world = _dist_world()

    def fn(device, segs):
        return _run_sites_part(device, traj, natoms, segs, n_blocks, n_frames,
                               site, nbins, rng, edges)

    _tp("blocks: sharding")
    outs, _ = _local_parts(fn, devices, parts, fetch_now=False)
    _tp("blocks: parts queued and finished on the device")
    st = outs[0]["st"]
    cnt = outs[0]["res"]
    for o in outs[1:]:                  # other local devices
        cnt = cnt + o["res"].to(cnt.device)
    extra = np.stack([sum(o["vol"] for o in outs),
                      sum(o["io"] for o in outs),
                      sum(o["compute"] for o in outs)])
    if world > 1:
        cnt = _allreduce(cnt)
        extra = _allreduce(torch.from_numpy(extra)).numpy()
    first = np.min(np.stack([o["first_launch"] for o in outs]), axis=0)

    # per-block volumes, summed in frame order like `res += r`
    vols, pos = [], 0
    for blk in blocks:
        v = 0.0
        for x in extra[0][pos:pos + len(blk)]:
            v = v + float(x)
        vols.append(np.array(np.float64(v)))
        pos += len(blk)
    den = None if den_fn is None else np.ascontiguousarray(den_fn(vols),
                                                           dtype=np.float64)
    # epilogue: the int32 counts come to the host in one D2H copy; the
    # float64 tensors the reference returns, their sum over the blocks and
    # the rdf are made from them by the threaded native pass
    # (rdfk_host_sites_finish: 3x fewer PCIe bytes than shipping float64, and
    # the results are ordinary numpy arrays the caller may keep)
    want_total = den is not None and n_blocks > 1
    with st.lock:
        if st.is_cuda:
            with torch.cuda.device(st.device):
                flat = cnt.reshape(-1)
                even = flat.numel() % 2 == 0
                cnt_host = _result_to_host(
                    st, flat.view(torch.int64) if even
                    else flat.to(torch.int64), st.compute_streams[0])
                cnt_host = cnt_host.view(np.int32) if even \
                    else cnt_host.astype(np.int32)
        else:
            cnt_host = cnt.numpy()
        cnt_host = np.ascontiguousarray(cnt_host).reshape(n_blocks, n)
        _tp("blocks: counts on the host")
        host, total, rdf_flat = _backend.host_sites_finish(
            cnt_host, nbins, den, want_total, Config.gather_threads)
        _tp("blocks: native epilogue")
    cell_off = np.asarray(site["cell_off"], dtype=np.int64)
    shapes = site["shapes"]

    def per_site_views(flat):
        dense = flat.reshape(total_cells, nbins)
        per_site = np.empty(len(shapes), dtype=object)
        for s_, (n1, n2) in enumerate(shapes):
            per_site[s_] = dense[cell_off[s_]:cell_off[s_ + 1]].reshape(
                n1, n2, nbins)
        return per_site

    def make_result(b, flat, vol):
        r = np.empty(2, dtype=object)
        r[0] = per_site_views(flat)
        r[1] = np.array(vol, dtype=np.float64)
        return r

    res = _finish(host, extra[0], extra[1], extra[2], first, blocks,
                  make_result)
    _tp("blocks: results assembled")
    if den is None:
        return res
    if not want_total:
        total = host[0]          # one block: the sum is the block itself
    return res, {"count": per_site_views(total),
                 "rdf": list(per_site_views(rdf_flat)), "den": den}


# --------------------------------------------------------------------------
# Contacts (per-frame time series instead of an accumulated histogram)
# --------------------------------------------------------------------------
def _run_contacts_part(device, traj, natoms, segs, n_blocks, n_frames, spec):
    st = _state(device)
    be = _backend
    width = spec["width"]
    out = {
        "hist": None,
        "vol": np.zeros(n_frames, dtype=np.float64),
        "io": np.zeros(n_frames, dtype=np.float64),
        "compute": np.zeros(n_frames, dtype=np.float64),
        "first_launch": np.full(n_blocks, np.inf),
    }
    with st.lock:
        cuda = st.is_cuda
        ctx = torch.cuda.device(device) if cuda else _nullctx()
        with ctx:
            # one row per frame of the run (frames of other devices stay 0).
            # User callables (CONTACT_DIST) get their distances batch by
            # batch: only the callable's value per (frame, reference) is kept,
            # never the [n_frames, n_pairs] distance table of the whole run.
            emit = spec["method"] == _cabi.CONTACT_DIST
            reduce_fn = spec.get("reduce") if emit else None
            keep = len(spec["refs"]) if reduce_fn is not None else width
            series = torch.zeros((n_frames, keep), dtype=torch.float64,
                                 device=device)
            series_host = np.zeros((n_frames, keep)) \
                if reduce_fn is not None else None
            refs = []
            for ref in spec["refs"]:
                refs.append((_to_dev_i32(ref["ia"], device),
                             _to_dev_i32(ref["ib"], device),
                             torch.as_tensor(ref["r0"], dtype=torch.float64,
                                             device=device),
                             len(ref["ia"]), ref["col"]))
            pipe = _Pipeline(st, traj, natoms)
            if emit:
                # keep batches of emitted distances bounded (256 MB)
                cap = max(1, (256 << 20) // max(8, 8 * width))
                pipe.batch_frames = max(1, min(pipe.batch_frames, cap))
            timers = []
            if cuda:
                st.compute_stream.wait_stream(torch.cuda.current_stream(device))
            for (b, pos0, frames) in segs:
                done = 0
                for batch in pipe.batches(frames):
                    B = len(batch)
                    xyz, dims, t_io, slot = pipe.fetch(batch)
                    lo_f = pos0 + done
                    out["io"][lo_f:lo_f + B] = t_io / B
                    out["first_launch"][b] = min(out["first_launch"][b],
                                                 time.time())
                    tm = _Timer(st)
                    tm.start()
                    stream = st.compute_stream.cuda_stream if cuda else 0
                    sctx = torch.cuda.stream(st.compute_stream) if cuda \
                        else _nullctx()
                    if reduce_fn is not None:
                        with sctx:
                            dest = torch.empty((B, width), dtype=torch.float64,
                                               device=device)
                    else:
                        dest = series[lo_f:lo_f + B]
                    for (ia, ib, r0, npairs, col) in refs:
                        be.contacts(xyz, B, natoms, ia, ib, r0, npairs,
                                    spec["method"], spec["radius"],
                                    spec["beta"], spec["lambda"],
                                    dest[:, col:], width, stream)
                    tm.stop()
                    pipe.mark_compute_done(slot)
                    if reduce_fn is not None:
                        with sctx:
                            dist_host = dest.to("cpu").numpy()  # synchronises
                        series_host[lo_f:lo_f + B] = reduce_fn(dist_host)
                    timers.append((tm, lo_f, B))
                    done += B
            for tm, lo_f, B in timers:
                out["compute"][lo_f:lo_f + B] = tm.seconds() / B
            if cuda:
                st.sync_compute()
            pipe.verify()       # cache hits still match the trajectory?
            if series_host is not None:
                series = torch.as_tensor(series_host, device=device)
            out["hist"] = series
    return out


def run_contacts_blocks(traj, natoms, spec, blocks, n_jobs):
    """Contacts over frame blocks.  ``spec``: ``refs`` = list of dicts with the
    contact pairs of one reference (``ia``, ``ib`` absolute atom indices,
    ``r0`` float64, ``col`` = first output column), ``method`` (rdfk.h
    RDFK_CONTACT_*), ``radius``, ``beta``, ``lambda``, ``width`` = columns per
    frame.  -> per-block tuples whose first element is the float64
    ``[n_frames_of_block, width]`` array of per-frame sums / distances."""
    devices = _backend.devices(n_jobs)
    n_blocks = len(blocks)
    parts, n_frames = _shard_for_this_rank(blocks, devices)

    def fn(device, segs):
        return _run_contacts_part(device, traj, natoms, segs, n_blocks,
                                  n_frames, spec)

    outs = _run_parts(fn, devices, parts)
    keep = len(spec["refs"]) if (spec["method"] == _cabi.CONTACT_DIST and
                                 spec.get("reduce") is not None) \
        else spec["width"]
    series = _combine(outs, devices[0], n_frames, keep, torch.float64)
    series = _allreduce(series)          # every frame was computed by one rank
    extra = np.stack([sum(o["vol"] for o in outs),
                      sum(o["io"] for o in outs),
                      sum(o["compute"] for o in outs)])
    if _dist_world() > 1:
        extra = _allreduce(torch.from_numpy(extra)).numpy()
    first = np.min(np.stack([o["first_launch"] for o in outs]), axis=0)
    series_np = series.cpu().numpy()
    res, pos = [], 0
    for b, blk in enumerate(blocks):
        n = len(blk)
        t_io = extra[1][pos:pos + n].copy()
        t_c = extra[2][pos:pos + n].copy()
        wait_end = first[b] if np.isfinite(first[b]) else time.time()
        res.append((series_np[pos:pos + n].copy(), t_io, t_c, 0.0, wait_end,
                    np.sum(t_io), np.sum(t_c)))
        pos += n
    return res


# --------------------------------------------------------------------------
# HydrogenBondAnalysis / capped_distance: ragged per-frame results
# --------------------------------------------------------------------------
def _run_hbonds_part(device, traj, natoms, segs, n_blocks, n_frames, spec):
    """All frames of one local device.  -> dict with ``rows``: list of
    (position of the batch's first frame in the run, frames in the batch,
    row int32, acc int32, dist float64, angle float64) host arrays."""
    st = _state(device)
    be = _backend
    out = {
        "rows": [],
        "io": np.zeros(n_frames, dtype=np.float64),
        "compute": np.zeros(n_frames, dtype=np.float64),
        "first_launch": np.full(n_blocks, np.inf),
    }
    don, hyd, acc = spec["donors"], spec["hydrogens"], spec["acceptors"]
    n_dh, n_acc = len(don), len(acc)
    with st.lock:
        cuda = st.is_cuda
        ctx = torch.cuda.device(device) if cuda else _nullctx()
        with ctx:
            d_d = _to_dev_i32(don, device)
            h_d = None if hyd is None else _to_dev_i32(hyd, device)
            a_d = _to_dev_i32(acc, device)
            total_d = torch.zeros(1, dtype=torch.int64, device=device)
            pipe = _Pipeline(st, traj, natoms)
            # keep the (frame, donor) row tables of one call bounded (256 MB)
            cap_f = max(1, (256 << 20) // max(56, 56 * n_dh))
            # (and the per-frame cell table of the acceptor grid: 128 KB)
            pipe.batch_frames = max(1, min(pipe.batch_frames, cap_f, 4096))
            ws = st.get_workspace(be.hbonds_workspace_bytes(
                pipe.batch_frames, n_dh, n_acc))
            capacity = 0
            bufs = None
            timers = []
            if cuda:
                st.compute_stream.wait_stream(torch.cuda.current_stream(device))
            for (b, pos0, frames) in segs:
                done = 0
                for batch in pipe.batches(frames):
                    B = len(batch)
                    xyz, dims, t_io, slot = pipe.fetch(batch)
                    kinds, box9, _ = _box_rows(dims, B)
                    lo_f = pos0 + done
                    out["io"][lo_f:lo_f + B] = t_io / B
                    box_d = _dev_const(st, box9, np.float32)
                    out["first_launch"][b] = min(out["first_launch"][b],
                                                 time.time())
                    tm = _Timer(st)
                    tm.start()
                    stream = st.compute_stream.cuda_stream if cuda else 0
                    sctx = torch.cuda.stream(st.compute_stream) if cuda \
                        else _nullctx()
                    for kind, lo, hi in _runs(kinds):
                        want = max(4096, 2 * (hi - lo) * n_dh)
                        while True:
                            if capacity < want:
                                capacity = int(want)
                                with sctx:
                                    bufs = (
                                        torch.empty(capacity, dtype=torch.int32,
                                                    device=device),
                                        torch.empty(capacity, dtype=torch.int32,
                                                    device=device),
                                        torch.empty(capacity,
                                                    dtype=torch.float64,
                                                    device=device),
                                        torch.empty(capacity,
                                                    dtype=torch.float64,
                                                    device=device))
                            be.hbonds(xyz[lo:hi], hi - lo, natoms, d_d, h_d,
                                      n_dh, a_d, n_acc, kind, box_d[lo:hi],
                                      spec["max_cutoff"], spec["min_cutoff"],
                                      spec["angle_cutoff"], capacity, bufs[0],
                                      bufs[1], bufs[2], bufs[3], total_d, ws,
                                      stream)
                            with sctx:
                                total = int(total_d.cpu()[0])   # synchronises
                            if total <= capacity:
                                break
                            want = total + total // 4       # ragged: grow, redo
                        with sctx:
                            # (copy=True: the stand-in CPU backend of the
                            # tests would otherwise alias the reused buffers)
                            got = tuple(t[:total].to("cpu", copy=True).numpy()
                                        for t in bufs)
                        out["rows"].append((lo_f + lo, hi - lo) + got)
                    tm.stop()
                    pipe.mark_compute_done(slot)
                    timers.append((tm, lo_f, B))
                    done += B
            for tm, lo_f, B in timers:
                out["compute"][lo_f:lo_f + B] = tm.seconds() / B
            if cuda:
                st.sync_compute()
            pipe.verify()       # cache hits still match the trajectory?
    return out


def run_hbonds_blocks(traj, natoms, spec, blocks, n_jobs):
    """Hydrogen-bond search over frame blocks.  ``spec``: ``donors`` /
    ``hydrogens`` (paired absolute atom indices; ``hydrogens`` None = plain
    capped_distance), ``acceptors``, ``max_cutoff``, ``min_cutoff`` (None =
    no lower bound), ``angle_cutoff`` (degrees).  -> per-block tuples whose
    first element is the column tuple ``(frame, k, acceptor, distance,
    angle)`` -- position of the frame in the run (int64), D-H row, position
    in ``acceptors``, float64 distance and angle -- ordered by (frame, k,
    acceptor)."""
    devices = _backend.devices(n_jobs)
    n_blocks = len(blocks)
    parts, n_frames = _shard_for_this_rank(blocks, devices)
    n_dh = max(len(spec["donors"]), 1)

    def fn(device, segs):
        return _run_hbonds_part(device, traj, natoms, segs, n_blocks, n_frames,
                                spec)

    outs = _run_parts(fn, devices, parts)
    chunks = [c for o in outs for c in o["rows"]]
    extra = np.stack([sum(o["io"] for o in outs),
                      sum(o["compute"] for o in outs)])
    if _dist_world() > 1:
        import torch.distributed as dist
        gathered = [None] * _dist_world()
        dist.all_gather_object(gathered, chunks)    # ragged: one exchange
        chunks = [c for part in gathered for c in part]
        extra = _allreduce(torch.from_numpy(extra)).numpy()
    chunks.sort(key=lambda c: c[0])                 # by first frame
    if chunks:
        row = np.concatenate([c[2] for c in chunks]).astype(np.int64)
        base = np.concatenate([np.full(len(c[2]), c[0], dtype=np.int64)
                               for c in chunks])
        frame = base + row // n_dh
        k = row % n_dh
        acc = np.concatenate([c[3] for c in chunks]).astype(np.int64)
        dist_ = np.concatenate([c[4] for c in chunks])
        ang = np.concatenate([c[5] for c in chunks])
    else:
        frame = k = acc = np.empty(0, dtype=np.int64)
        dist_ = ang = np.empty(0, dtype=np.float64)
    first = np.min(np.stack([o["first_launch"] for o in outs]), axis=0)
    res, pos = [], 0
    starts = np.searchsorted(frame, np.arange(n_frames + 1), side="left")
    for b, blk in enumerate(blocks):
        n = len(blk)
        t_io = extra[0][pos:pos + n].copy()
        t_c = extra[1][pos:pos + n].copy()
        wait_end = first[b] if np.isfinite(first[b]) else time.time()
        sl = slice(starts[pos], starts[pos + n])
        cols = (frame[sl], k[sl], acc[sl], dist_[sl], ang[sl])
        res.append((cols, t_io, t_c, 0.0, wait_end, np.sum(t_io), np.sum(t_c)))
        pos += n
    return res
