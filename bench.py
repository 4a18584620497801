#!/usr/bin/env python
"""bench.py -- InterRDF pair-distances/sec on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...        # CPU arm (oracle port)

Headline workload: BASELINE.json configs[1] -- InterRDF all-atom on a synthetic
100,000-atom orthorhombic box (L = 100 A), nbins = 200, range 0-15 A.  One
*step* is one ``InterRDF.run()`` over ``--frames-per-step`` frames per GPU
(default 256 frames = 307 MB of coordinates per GPU, larger than the 126 MB
L2).  Frames are sharded over ranks by the engine; per-rank int64 histograms
are combined by one all-reduce inside ``run()`` (weak scaling).

One JSON line on rank 0:
  value          pair-distances/s, whole job, frames resident in HBM (device
                 frame cache hit, validated by a hash of every host byte that
                 runs beside the kernels), barrier + synchronize on both
                 sides, max over ranks
  e2e            the same through the public API with HOST buffers: each step
                 copies its frames host->device and reads the histogram back
  roofline       the pair kernel against the FP32 peak measured live, on the
                 pairs it really EVALUATES (frac); the algorithmic 20-FLOP
                 accounting sits beside it under `algorithmic`
  configs        short timed runs of BASELINE configs[0], [2], [3], [4]
                 (frame-sharded over the N ranks) with an in-bench oracle
                 check each
  parity_nranks  a fixed seeded trajectory through run() on all N ranks
                 against the committed oracle golden (tests/golden/
                 nranks_counts.npz); false fails the run
  strong         256 frames in total split over the N ranks
  cpu_baseline   the oracle port on all host cores (rank 0, N = 1 only)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_ATOMS = 100_000
NBINS = 200
RANGE = (0.0, 15.0)
BOX_L = 100.0
FLOP_PER_PAIR = 20.0          # SURVEY.md 8(d), orthorhombic, algorithmic
FLOP_PER_EVALUATED_PAIR = 8.0  # 3 sub + 3 mul + 2 add: what the kernel does
WORKLOAD = ("BASELINE configs[1]: InterRDF all-atom, synthetic 100k-atom "
            "orthorhombic box, nbins=200, range 0-15 A")
STRONG_FRAMES = 256


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=256,
                    help="frames per GPU per step")
    ap.add_argument("--n-atoms", type=int, default=N_ATOMS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the configs / strong / parity_nranks legs")
    return ap.parse_args()


def host_cores():
    """cores this process may run on -- NOT omp_get_max_threads(): under
    torch.distributed.run OMP_NUM_THREADS is 1 and the CPU arm of round 1 ran
    on one core"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bench_config(n_atoms, fps):
    """the `config` both arms print (the driver compares them)"""
    return {"workload": WORKLOAD, "n_atoms": n_atoms, "nbins": NBINS,
            "range": list(RANGE), "frames_per_step_per_gpu": fps,
            "l2": "inputs per step (%.0f MB/GPU) exceed the 126 MB L2"
                  % (fps * n_atoms * 12 / 1e6),
            "parallelism": "frames sharded over the ranks, one all-reduce "
                           "per run()"}


def synth_positions(n_frames, n_atoms, seed, only=None):
    """uniform ideal gas in [0, L).  Every frame has its own seed, so a rank
    can generate just the frames of its shard (`only`); the others stay
    untouched memory it never reads."""
    pos = np.empty((n_frames, n_atoms, 3), dtype=np.float32)
    hi = np.nextafter(np.float32(BOX_L), np.float32(0))
    for f in (range(n_frames) if only is None else only):
        rng = np.random.Generator(np.random.PCG64([seed, f]))
        pos[f] = (rng.random((n_atoms, 3)) * BOX_L).astype(np.float32)
        np.clip(pos[f], 0, hi, out=pos[f])
    return pos


# ---------------------------------------------------------------------------
# clocks sampler (recipe line of B200_PROFILING.md)
# ---------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [],
                    "note": "nvidia-smi unavailable"}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for name, val in zip(names, p[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ---------------------------------------------------------------------------
# MDAnalysis' capped_distance (the call behind pmda/rdf.py:123) picks its grid
# search for this workload (N1*N2 >= 1e8); results are those of
# distance_array + np.histogram, which is what the oracle restates.  The CPU
# arm therefore times the oracle's CELL-LIST variant (same integers, verified
# against the full pair matrix in tests/test_oracle.py) -- the faster of the
# two CPU algorithms, one OpenMP loop with the histogram in C: a STRONGER CPU
# baseline than pmda itself (process per block, per-frame Python, a float64
# distance temporary and np.histogram) -- and reports the full-matrix rate
# next to it.
def cpu_sample(n_atoms, n_rows, threads, seed=99, cells=True):
    """one bounded sample: n_rows x n_atoms pairs of one frame -> (s, pairs)"""
    from oracle import c_oracle
    pos = synth_positions(1, n_atoms, seed)[0]
    dims = [BOX_L, BOX_L, BOX_L, 90.0, 90.0, 90.0]
    t0 = time.perf_counter()
    c_oracle.rdf_single_frame(pos[:n_rows], pos, dims, NBINS, RANGE,
                              nthreads=threads, cells=cells)
    return time.perf_counter() - t0, float(n_rows) * n_atoms


def cpu_rows_for(threads, n_atoms, seconds=4.0, cells=True):
    """rows of the pair matrix that take ~`seconds` (35 Mpairs/s/thread for
    the full matrix, ~12x that through the cell list at this density)"""
    rate = 35e6 * (12.0 if cells else 1.0)
    rows = int(seconds * rate * threads / n_atoms)
    return max(64, min(n_atoms, rows))


CPU_NOTE = ("stronger than pmda: one OpenMP C loop over rows with the "
            "histogram in C, through a cell list; pmda itself runs a process "
            "per block with per-frame Python, a float64 distance temporary "
            "and np.histogram (that shape: configs.cfg1.cpu_pmda_shape)")


def cpu_baseline_dict(n_atoms, threads, seconds):
    rows = cpu_rows_for(threads, n_atoms, seconds=seconds)
    cpu_sample(n_atoms, min(rows, 2048), threads)               # warm
    dt, p = cpu_sample(n_atoms, rows, threads)
    dt2, p2 = cpu_sample(n_atoms, rows, threads)
    rows_b = cpu_rows_for(threads, n_atoms, seconds=2.0, cells=False)
    dtb, pb = cpu_sample(n_atoms, rows_b, threads, cells=False)
    return {"value": (p + p2) / (dt + dt2), "unit": "pairs/s",
            "cores": threads, "kind": "port",
            "sample": "2 x %d of %d rows of one frame's pair matrix (%.3g "
                      "algorithmic pairs) through the cell-list oracle, OpenMP "
                      "over rows" % (rows, n_atoms, p + p2),
            "full_pair_matrix_value": pb / dtb,
            "full_pair_matrix_sample": "%d rows, every pair evaluated "
                                       "(distance_array semantics)" % rows_b,
            "note": CPU_NOTE}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_cores()
    rows = cpu_rows_for(threads, args.n_atoms)
    for _ in range(args.warmup):
        cpu_sample(args.n_atoms, rows, threads)
    t = 0.0
    pairs = 0.0
    for _ in range(args.steps):
        dt, p = cpu_sample(args.n_atoms, rows, threads)
        t += dt
        pairs += p
    value = pairs / t
    rows_b = cpu_rows_for(threads, args.n_atoms, seconds=2.0, cells=False)
    dtb, pb = cpu_sample(args.n_atoms, rows_b, threads, cells=False)
    sample = ("%d of %d rows of the pair matrix of one frame per step "
              "(%.3g algorithmic pairs/step) through the cell-list oracle, "
              "OpenMP over rows on %d threads" %
              (rows, args.n_atoms, float(rows) * args.n_atoms, threads))
    line = {
        "impl": "reference", "metric": "InterRDF pair-distances/sec",
        "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "frames_per_s": value / (float(args.n_atoms) * args.n_atoms),
        "config": bench_config(args.n_atoms, args.frames_per_step),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads,
                         "kind": "port", "sample": sample,
                         "full_pair_matrix_value": pb / dtb,
                         "note": CPU_NOTE},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "cores": threads,
        "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
        "note": ("MDAnalysis/dask are not installable here; this is the C "
                 "oracle restating capped_distance + np.histogram on all host "
                 "cores (sched_getaffinity, whatever OMP_NUM_THREADS says), "
                 "through its cell list (MDAnalysis dispatches to a grid "
                 "search at this size); full_pair_matrix_value is the "
                 "distance_array rate"),
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
class Job(object):
    """torch.distributed plumbing of one bench process"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(
                "nccl", device_id=torch.device("cuda", self.local_rank))
        if self.world != args.gpus and self.rank == 0:
            print("warning: --gpus %d but WORLD_SIZE %d"
                  % (args.gpus, self.world), file=sys.stderr)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def min(self, x):
        return self._reduce(x, self.dist.ReduceOp.MIN)

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def my_frames(self, n_frames):
        """the frames the engine gives this rank when `n_frames` are cut into
        `world` blocks"""
        from pmda_b200 import engine
        from pmda_b200.util import make_balanced_slices
        blocks = [range(sl.start, sl.stop, sl.step) for sl in
                  make_balanced_slices(n_frames, self.world, start=0,
                                       stop=n_frames, step=1)]
        return [f for (_, _, fr) in
                engine._frame_shards(blocks, self.world)[self.rank]
                for f in fr]

    def timed(self, step, k):
        """k calls of step() between barrier + synchronize pairs -> (seconds
        max over ranks, kernel seconds of this rank's timers, launches, last
        result)"""
        from pmda_b200 import _cabi
        self.barrier()
        l0 = _cabi.launch_count()
        t0 = time.perf_counter()
        kernel_s = 0.0
        last = None
        for _ in range(k):
            last = step()
            kernel_s += float(last.timing.compute.sum())
        self.torch.cuda.synchronize()
        t1 = time.perf_counter()
        self.barrier()
        return (self.max(t1 - t0), kernel_s, _cabi.launch_count() - l0, last)

    def resident_and_e2e(self, step, k, warm):
        """(t_resident, t_e2e, kernel_s, launches, last) of k steps each"""
        from pmda_b200 import engine
        engine.Config.cache_bytes = 64 << 30
        engine.clear_cache()
        for _ in range(max(warm, 1)):
            step()
        t_dev, kernel_s, launches, last = self.timed(step, k)
        engine.Config.cache_bytes = 0
        engine.clear_cache()
        for _ in range(max(warm, 1)):
            step()
        t_e2e, _, _, last_e2e = self.timed(step, k)
        engine.Config.cache_bytes = 64 << 30
        return t_dev, t_e2e, kernel_s, launches, last, last_e2e


def last_rdf_stats(job):
    """librdfk's counters of the last rdfk_rdf_hist call on this rank"""
    from pmda_b200 import _cabi, engine
    st = engine._state(job.torch.device("cuda", job.local_rank))
    ws = st.workspaces[0]
    if ws is None:
        return None
    ws = ws[(-ws.data_ptr()) % 256:]
    return _cabi.rdf_stats(ws, st.compute_streams[0].cuda_stream)


def committed_ncu():
    """numbers of the committed ncu capture of the pair kernel (profiles/)"""
    for name in ("r2_traffic.json", "r2_before_traffic.json", "r1h_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                d = json.load(fh)
            d["file"] = "profiles/" + name
            return d
        except Exception:
            continue
    return None


# -- the other BASELINE configurations ---------------------------------------
# (configs[2] at the trajectory length BASELINE.json names: with 400 frames the
# fixed cost of a run() hid a third of the throughput, 72 000 against 112 000
# frames/s end to end)
# configs[3] / [4]: 384 MB / 768 MB of frames per GPU, enough for the first
# upload (the only one no kernel hides) not to dominate a run
CONFIG_FRAMES = {1: ("total", 100), 3: ("per_gpu", 2000), 4: ("per_gpu", 128),
                 5: ("per_gpu", 64)}
CONFIG_NAMES = {
    1: "BASELINE configs[0]: InterRDF O-O, 3,000-water box, 100 frames, "
       "nbins=75, range 0-15 A, orthorhombic",
    3: "BASELINE configs[2]: InterRDF_s, 64 ion-water site pairs (1 x 256), "
       "50k-atom box, 2000 frames, nbins=75",
    4: "BASELINE configs[3]: InterRDF, 250k-atom triclinic box "
       "(152.3 A, 60/60/90), nbins=75, frame-sharded",
    5: "BASELINE configs[4]: InterRDF 1k x 1M membrane-like box, nbins=75, "
       "frame-sharded, one histogram all-reduce per run()",
}


def oracle_check(job, cfg, u, kw):
    """in-bench parity: the GPU path through run() against the oracle on a
    bounded piece of the same trajectory (frame 0 lives on rank 0)"""
    from pmda_b200.rdf import InterRDF, InterRDF_s
    from oracle import c_oracle
    threads = host_cores()
    ts_pos = u.trajectory.positions
    dims = u.trajectory.dimensions[0]
    if cfg == 3:
        r = InterRDF_s(u, kw["ags"], nbins=kw["nbins"],
                       range=kw["range"]).run(stop=1, n_blocks=1)
        if job.rank != 0:
            return True, "frame 0, all 64 site pairs"
        pairs = [(ts_pos[0][a.indices], ts_pos[0][b.indices])
                 for a, b in kw["ags"]]
        want, _ = c_oracle.rdf_s_single_frame(pairs, dims, kw["nbins"],
                                              kw["range"])
        ok = all(np.array_equal(r.count[i], want[i])
                 for i in range(len(want)))
        return ok, "frame 0, all 64 site pairs"
    g1, g2 = kw["g1"], kw["g2"]
    what = "frame 0, full pair matrix"
    if cfg == 4:      # 27-image oracle: a row subsample
        pick = np.sort(np.random.default_rng(4).choice(len(g1), 48,
                                                       replace=False))
        g1 = g1[pick]
        what = "frame 0, 48 random rows x 250k atoms"
    r = InterRDF(g1, g2, nbins=kw["nbins"], range=kw["range"]).run(
        stop=1, n_blocks=1)
    if job.rank != 0:
        return True, what
    want = np.zeros(kw["nbins"], dtype=np.int64)
    c_oracle.rdf_single_frame(ts_pos[0][g1.indices], ts_pos[0][g2.indices],
                              dims, kw["nbins"], kw["range"], count=want,
                              nthreads=threads, cells=True)
    return bool(np.array_equal(r.count, want)), what


def run_config(job, cfg, reps=None, warm=3, cpu_shape=False):
    """frames/s of one BASELINE configuration, resident and end to end,
    frame-sharded over the ranks"""
    if reps is None:      # millisecond runs: more of them
        reps = 20 if cfg in (1, 3) else 4
    from pmda_b200 import engine, synthetic
    from pmda_b200.rdf import InterRDF, InterRDF_s
    mode, n = CONFIG_FRAMES[cfg]
    n_frames = n if mode == "total" else n * job.world
    mine = job.my_frames(n_frames)
    only = sorted(set(mine) | ({0} if job.rank == 0 else set()))
    u, kw = synthetic.sharded_config(cfg, n_frames, only=only)
    if cfg == 3:
        def make():
            return InterRDF_s(u, kw["ags"], nbins=kw["nbins"],
                              range=kw["range"])
        pairs = float(sum(len(a) * len(b) for a, b in kw["ags"]))
    else:
        def make():
            return InterRDF(kw["g1"], kw["g2"], nbins=kw["nbins"],
                            range=kw["range"])
        pairs = float(len(kw["g1"])) * len(kw["g2"])

    def step():
        return make().run(n_blocks=job.world)

    t_dev, t_e2e, kernel_s, launches, last, last_e2e = \
        job.resident_and_e2e(step, reps, warm)
    if cfg == 3:
        same = all(np.array_equal(a, b)
                   for a, b in zip(last.count, last_e2e.count))
    else:
        same = bool(np.array_equal(last.count, last_e2e.count))
    out = {
        "workload": CONFIG_NAMES[cfg], "frames": n_frames,
        "n_gpus": job.world, "pairs_per_frame": pairs,
        "resident_frames_per_s": n_frames * reps / t_dev,
        "e2e_frames_per_s": n_frames * reps / t_e2e,
        "resident_pairs_per_s": n_frames * reps * pairs / t_dev,
        "e2e_pairs_per_s": n_frames * reps * pairs / t_e2e,
        "run_ms_resident": 1e3 * t_dev / reps, "run_ms_e2e": 1e3 * t_e2e / reps,
        "kernel_frames_per_s_per_gpu":
            n_frames * reps / kernel_s if kernel_s > 0 else None,
        "h2d_bytes_per_run_per_gpu": len(mine) * u.atoms.n_atoms * 12,
        "resident_equals_e2e": same,
    }
    if cfg != 3:
        stats = last_rdf_stats(job)
        if stats is not None and stats["tile_items"]:
            ncl = -(-max(len(kw["g1"]), len(kw["g2"])) // 128)
            frames_last = max(1, stats["tile_items"] // ncl)
            out["evaluated_fraction"] = \
                stats["tile_pairs"] / float(frames_last) / pairs
            out["fp64_rechecked_pairs_per_frame"] = \
                stats["slow_pairs"] / float(frames_last)
    ok, what = oracle_check(job, cfg, u, kw)
    ok = job.min(1.0 if (ok and same) else 0.0) == 1.0
    out["parity"] = ok
    out["parity_check"] = what + " vs the C oracle; resident == e2e counts"
    if cpu_shape and cfg == 1 and job.rank == 0 and job.world == 1:
        out["cpu_pmda_shape"] = cpu_pmda_shape(u, kw)
    engine.clear_cache()
    del u, kw
    return out


def cpu_pmda_shape(u, kw):
    """configs[0] as BASELINE.json asks for it on the CPU: n_jobs=4, one
    process per block, per-frame distance_array + mask + np.histogram
    (oracle/pmda_cpu.py) -- on a bounded sample of 3 frames per worker"""
    from oracle import pmda_cpu
    dims = u.trajectory.dimensions[0]
    ix = kw["g1"].indices
    out = {}
    for jobs in sorted({4, min(host_cores(), 16)}):
        n_frames = min(3 * jobs, len(u.trajectory.positions))
        pos = u.trajectory.positions[:n_frames]
        r = pmda_cpu.interrdf_process_per_block(
            pos, dims, ix, ix, kw["nbins"], kw["range"], n_jobs=jobs)
        out["n_jobs_%d" % jobs] = {
            "frames_per_s": n_frames / r["seconds"],
            "frames_per_s_with_worker_startup":
                n_frames / r["seconds_with_startup"],
            "sample": "%d frames, %d blocks, one spawned process per block"
                      % (n_frames, r["n_blocks"])}
    return out


def parity_nranks(job):
    """The invariance pmda/test/test_rdf.py:76-93 pins (same result for every
    number of blocks / workers): a fixed seeded 64-frame trajectory, ortho and
    triclinic, InterRDF and InterRDF_s, through run() on all ranks, against
    the committed oracle golden."""
    from pmda_b200 import engine
    from pmda_b200.rdf import InterRDF, InterRDF_s
    from pmda_b200.universe import Universe
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_nranks_golden as mk
    gold = np.load(os.path.join(ROOT, "tests", "golden", "nranks_counts.npz"))
    detail = {}
    ok_all = True
    for name in sorted(mk.CASES):
        mine = job.my_frames(mk.N_FRAMES)
        pos, dims = mk.make_positions(name, mine)
        _, _, nbins, rng = mk.CASES[name]
        u = Universe(pos, dims)
        for nb in sorted({job.world, 2 * job.world + 1}):
            # (blocks != ranks: blocks then straddle the ranks' frame shards,
            # which are cut from the concatenated frame list)
            r = InterRDF(u.atoms, u.atoms, nbins=nbins, range=rng).run(
                n_blocks=nb)
            ok = bool(np.array_equal(r.count, gold[name + "__total"]))
            detail["%s_n_blocks_%d" % (name, nb)] = ok
            ok_all = ok_all and ok
        if name == "ortho":
            sites = mk.site_indices()
            ags = [[u.atoms[a], u.atoms[b]] for a, b in sites]
            s = InterRDF_s(u, ags, nbins=nbins, range=rng).run(
                n_blocks=job.world)
            got = np.stack([c[0] for c in s.count])
            ok = bool(np.array_equal(got, gold["sites__total"]))
            detail["sites_n_blocks_%d" % job.world] = ok
            ok_all = ok_all and ok
        engine.clear_cache()
    ok_all = job.min(1.0 if ok_all else 0.0) == 1.0
    return ok_all, detail


def run_b200(args):
    job = Job(args)
    torch = job.torch
    from pmda_b200 import _cabi, engine
    from pmda_b200.rdf import InterRDF
    from pmda_b200.universe import Universe
    world, rank = job.world, job.rank

    fps = args.frames_per_step
    n_frames = fps * world                       # weak scaling
    mine = job.my_frames(n_frames)
    pos = synth_positions(n_frames, args.n_atoms, 20261019 + 2, only=mine)
    u = Universe(pos, [BOX_L, BOX_L, BOX_L, 90.0, 90.0, 90.0])
    n1 = n2 = args.n_atoms
    pairs_per_step = float(n_frames) * n1 * n2

    def step():
        r = InterRDF(u.atoms, u.atoms, nbins=NBINS, range=RANGE)
        r.run(n_blocks=world)
        return r

    # ---- device-resident arm (value) --------------------------------------
    engine.Config.cache_bytes = 64 << 30
    engine.clear_cache()
    for _ in range(max(args.warmup, 1)):
        step()
    sampler = ClockSampler(job.local_rank)
    if rank == 0:
        sampler.start()
    t_dev, kernel_s, launches, last = job.timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = pairs_per_step * args.steps / t_dev
    launches_all = job.sum(float(launches))
    count_dev = last.count.copy()
    stats = last_rdf_stats(job)

    # ---- end-to-end arm (e2e): host buffers, H2D + D2H every step ---------
    engine.Config.cache_bytes = 0
    engine.clear_cache()
    for _ in range(max(args.warmup, 1)):
        step()
    t_e2e, _, _, last_e2e = job.timed(step, args.steps)
    e2e_value = pairs_per_step * args.steps / t_e2e
    assert np.array_equal(count_dev, last_e2e.count)
    h2d = fps * args.n_atoms * 12 + fps * 9 * 4      # per GPU per step
    d2h = world * NBINS * 8
    engine.Config.cache_bytes = 64 << 30

    # ---- strong scaling: a fixed 256-frame trajectory split over N ranks ---
    strong = None
    if not args.no_configs:
        if world == 1 and fps == STRONG_FRAMES:
            strong = {"frames_total": STRONG_FRAMES, "n_gpus": 1,
                      "resident_frames_per_s": n_frames * args.steps / t_dev,
                      "e2e_frames_per_s": n_frames * args.steps / t_e2e,
                      "run_ms_resident": 1e3 * t_dev / args.steps,
                      "run_ms_e2e": 1e3 * t_e2e / args.steps,
                      "note": "N = 1: the headline run itself"}
        else:
            del u, pos
            engine.clear_cache()
            mine_s = job.my_frames(STRONG_FRAMES)
            pos_s = synth_positions(STRONG_FRAMES, args.n_atoms,
                                    20261019 + 2, only=mine_s)
            us = Universe(pos_s, [BOX_L, BOX_L, BOX_L, 90.0, 90.0, 90.0])

            def step_s():
                return InterRDF(us.atoms, us.atoms, nbins=NBINS,
                                range=RANGE).run(n_blocks=world)
            k = max(4, args.steps)
            ts_dev, ts_e2e, _, _, _, _ = job.resident_and_e2e(step_s, k, 2)
            strong = {"frames_total": STRONG_FRAMES, "n_gpus": world,
                      "resident_frames_per_s": STRONG_FRAMES * k / ts_dev,
                      "e2e_frames_per_s": STRONG_FRAMES * k / ts_e2e,
                      "run_ms_resident": 1e3 * ts_dev / k,
                      "run_ms_e2e": 1e3 * ts_e2e / k,
                      "note": "same 256 frames at every N (the reference's "
                              "own decomposition, pmda/parallel.py:365-392)"}
            del us, pos_s
            engine.clear_cache()

    # ---- the other BASELINE configurations, N-rank parity ------------------
    configs = None
    nranks_ok, nranks_detail = None, None
    if not args.no_configs:
        configs = {}
        for cfg in (1, 3, 4, 5):
            configs["cfg%d" % cfg] = run_config(
                job, cfg, cpu_shape=not args.no_cpu_baseline)
        nranks_ok, nranks_detail = parity_nranks(job)

    if rank != 0:
        if world > 1:
            job.dist.destroy_process_group()
        return 0 if nranks_ok in (None, True) else 1

    # ---- roofline of the dominant kernel (rank 0) --------------------------
    # rdfk_rdf_hist = init + stage + params + tables + sort (4) + pair_prune;
    # the CUDA events of engine._Timer bracket the whole call on the compute
    # stream, and the pair kernel is ~97 % of it (profiles/): its launch
    # duration is taken as kernel_s / calls
    peak = _cabi.measure_fp32_peak()
    # timing.compute holds every frame's kernel time (each frame ran on one
    # rank), so kernel_s sums over ranks and work / kernel_s is a per-GPU rate
    frames_rank = n_frames * args.steps
    ncl = -(-n1 // 128)
    frames_last = max(1, stats["tile_items"] // ncl)
    eval_pairs_per_frame = stats["tile_pairs"] / float(frames_last)
    eval_frac = eval_pairs_per_frame / (float(n1) * n2)
    eval_rate = eval_pairs_per_frame * frames_rank / kernel_s
    achieved = eval_rate * FLOP_PER_EVALUATED_PAIR / 1e12
    alg_flop = FLOP_PER_PAIR * frames_rank * float(n1) * n2
    alg_tflops = alg_flop / kernel_s / 1e12
    ncu = committed_ncu()
    traffic = None
    if ncu and args.n_atoms == N_ATOMS:
        traffic = ncu["dram_bytes_per_frame"] * frames_last
    hbm_peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm_peak = json.load(fh).get("hbm_gbs")
    except Exception:
        pass
    hbm_src = "MEASURED_PEAKS.json" if hbm_peak else "fallback 6650 GB/s"
    hbm_peak = hbm_peak or 6650.0
    alg_bytes = 12.0 * n1 * frames_rank           # g1 == g2: read once
    roofline = {
        "bound": "fp32-issue", "kernel": "pair_prune_kernel<ORTHO>",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak,
        "frac_definition": "evaluated pairs/s x 8 FLOP (3 sub + 3 mul + 2 "
                           "add: the arithmetic the kernel performs per "
                           "evaluated pair) / FP32 peak measured live",
        "issue_slot_frac": (ncu or {}).get("issue_slot_frac"),
        "issue_slot_source": (ncu or {}).get("file"),
        # the resource that is busiest in the committed capture: the kernel
        # queues, looks up and bins its in-range pairs through shared memory
        "shared_pipe_frac": (ncu or {}).get("shared_pipe_frac"),
        "binding_resource": "shared-memory pipe (shared_pipe_frac) and issue "
                            "slots (issue_slot_frac): neither HBM nor the FMA "
                            "pipe bounds this kernel (DESIGN.md section 4)",
        "traffic": traffic,
        "traffic_source": "%s: dram__bytes_read+write of one ncu --set full "
                          "capture, per frame x frames per launch"
                          % ((ncu or {}).get("file"),),
        "peak_source": "FFMA loop measured live (rdfk_measure_fp32_peak); "
                       "MEASURED_PEAKS.json has no FP32 figure",
        "launch_ms": 1e3 * kernel_s / args.steps / world /
                     max(1.0, fps / float(frames_last)),
        "frames_per_launch": frames_last,
        "evaluated": {
            "pairs_last_launch": stats["tile_pairs"],
            "fraction_of_pair_matrix": eval_frac,
            "pairs_per_s": eval_rate,
            "flop_per_evaluated_pair": FLOP_PER_EVALUATED_PAIR,
        },
        "algorithmic": {
            "flop_per_pair": FLOP_PER_PAIR,
            "tflops_equivalent": alg_tflops,
            "algorithmic_speedup_vs_fp32_peak": alg_tflops / peak,
            "note": "SURVEY 8d accounting (20 FLOP x N1 x N2 per frame) / "
                    "kernel time; exceeds the peak because exact pruning "
                    "never evaluates 1 - fraction_of_pair_matrix of the "
                    "pairs: a speed-up, not a utilisation",
        },
        "fp64_rechecked_pairs_last_launch": stats["slow_pairs"],
        "hbm": {"algorithmic_GBps": alg_bytes / kernel_s / 1e9,
                "peak_GBps": hbm_peak, "peak_source": hbm_src,
                "frac": alg_bytes / kernel_s / 1e9 / hbm_peak},
    }

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_dict(args.n_atoms, host_cores(), 6.0)

    line = {
        "metric": "InterRDF pair-distances/sec", "value": value,
        "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64", "data": "synthetic",
        "frames_per_s": n_frames * args.steps / t_dev,
        "config": bench_config(args.n_atoms, fps),
        "e2e": {"value": e2e_value, "unit": "pairs/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * t_e2e / args.steps,
                "frames_per_s": n_frames * args.steps / t_e2e},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "strong": strong,
        "configs": configs,
        "parity_nranks": nranks_ok,
        "parity_nranks_detail": nranks_detail,
        "count_sum": int(count_dev.sum()),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        job.dist.destroy_process_group()
    if nranks_ok is False:
        print("parity_nranks FAILED: %r" % (nranks_detail,), file=sys.stderr)
        return 1
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
