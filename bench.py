#!/usr/bin/env python
"""bench.py -- InterRDF pair-distances/sec on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...        # CPU arm (oracle port)

Workload: BASELINE.json configs[1] -- InterRDF all-atom on a synthetic
100,000-atom orthorhombic box (L = 100 A), nbins = 200, range 0-15 A.  One
*step* is one ``InterRDF.run()`` over a block of ``--frames-per-step`` frames
per GPU (default 256 frames = 307 MB of coordinates per GPU, larger than the
126 MB L2).  Frames are sharded over ranks by the engine; per-rank int64
histograms are combined by one all-reduce inside ``run()`` (weak scaling:
per-GPU work is fixed).

Printed keys (one JSON line, rank 0):
  value        pair-distances/s, whole job, inputs already resident in HBM
               (device frame cache hit), timed between barrier+synchronize
               pairs, max over ranks
  e2e          same metric through the public API with HOST buffers: every
               step copies its frames host->device (pinned) and reads the
               histogram back
  roofline     the tile kernel against the FP32 FFMA peak measured live
  cpu_baseline the oracle port on the host cores (rank 0, N = 1 only)
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
FLOP_PER_PAIR = 20.0          # SURVEY.md 8(d), orthorhombic
WORKLOAD = ("BASELINE configs[1]: InterRDF all-atom, synthetic 100k-atom "
            "orthorhombic box, nbins=200, range 0-15 A")


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
    return ap.parse_args()


def synth_positions(n_frames, n_atoms, seed, only=None):
    """uniform ideal gas in [0, L) (synthetic.config2 without the Universe).
    Every frame has its own seed, so a rank can generate just the frames of
    its shard (`only`); the others stay untouched memory it never reads."""
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
# two CPU algorithms -- and reports the full-matrix rate next to it.
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
                                       "(distance_array semantics)" % rows_b}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import c_oracle
    threads = c_oracle.max_threads()
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
              "OpenMP over rows" %
              (rows, args.n_atoms, float(rows) * args.n_atoms))
    line = {
        "impl": "reference", "metric": "InterRDF pair-distances/sec",
        "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "frames_per_s": value / (float(args.n_atoms) * args.n_atoms),
        "config": {"workload": WORKLOAD, "n_atoms": args.n_atoms,
                   "nbins": NBINS, "range": list(RANGE)},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads,
                         "kind": "port", "sample": sample,
                         "full_pair_matrix_value": pb / dtb},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": ("MDAnalysis/dask are not installable here; this is the C "
                 "oracle restating capped_distance + np.histogram on all host "
                 "threads, through its cell list (MDAnalysis dispatches to a "
                 "grid search at this size); full_pair_matrix_value is the "
                 "distance_array rate"),
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from pmda_b200 import _cabi, engine
    from pmda_b200.rdf import InterRDF
    from pmda_b200.universe import Universe

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda",
                                                              local_rank))
    if world != args.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE %d" % (args.gpus, world),
              file=sys.stderr)

    fps = args.frames_per_step
    n_frames = fps * world                       # weak scaling
    # the engine gives rank r the r-th of `world` balanced frame blocks
    from pmda_b200.util import make_balanced_slices
    blocks = [range(sl.start, sl.stop, sl.step) for sl in
              make_balanced_slices(n_frames, world, start=0, stop=n_frames,
                                   step=1)]
    mine = [f for (_, _, fr) in engine._frame_shards(blocks, world)[rank]
            for f in fr]
    pos = synth_positions(n_frames, args.n_atoms, 20261019 + 2, only=mine)
    u = Universe(pos, [BOX_L, BOX_L, BOX_L, 90.0, 90.0, 90.0])
    n1 = n2 = args.n_atoms
    pairs_per_step = float(n_frames) * n1 * n2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def step():
        r = InterRDF(u.atoms, u.atoms, nbins=NBINS, range=RANGE)
        r.run(n_blocks=world)
        return r

    def timed(k):
        barrier()
        l0 = _cabi.launch_count()
        t0 = time.perf_counter()
        kernel_s = 0.0
        last = None
        for _ in range(k):
            last = step()
            kernel_s += float(last.timing.compute.sum())
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        return (max_over_ranks(t1 - t0), kernel_s,
                _cabi.launch_count() - l0, last)

    # ---- device-resident arm (value) --------------------------------------
    engine.Config.cache_bytes = 64 << 30
    engine.clear_cache()
    for _ in range(max(args.warmup, 1)):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_dev, kernel_s, launches, last = timed(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = pairs_per_step * args.steps / t_dev
    launches_all = sum_over_ranks(float(launches))
    count_dev = last.count.copy()

    # ---- end-to-end arm (e2e): host buffers, H2D + D2H every step ---------
    engine.Config.cache_bytes = 0
    engine.clear_cache()
    for _ in range(max(args.warmup, 1)):
        step()
    t_e2e, _, _, last_e2e = timed(args.steps)
    e2e_value = pairs_per_step * args.steps / t_e2e
    assert np.array_equal(count_dev, last_e2e.count)
    h2d = fps * args.n_atoms * 12 + fps * 9 * 4      # per GPU per step
    d2h = world * NBINS * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (rank 0) --------------------------
    # rdfk_rdf_hist = init + stage + params + tables + sort (4) + pair_prune;
    # the CUDA events of engine._Timer bracket the whole call on the compute
    # stream, and the pair kernel is ~97 % of it (profiles/): its launch
    # duration is taken as kernel_s / calls
    peak = _cabi.measure_fp32_peak()
    # timing.compute holds every frame's kernel time (each frame ran on one
    # rank), so kernel_s sums over ranks and work/kernel_s is the per-GPU rate
    frames_all = n_frames * args.steps
    alg_flop = FLOP_PER_PAIR * frames_all * float(n1) * n2
    achieved = alg_flop / kernel_s / 1e12
    st = engine._state(torch.device("cuda", local_rank))
    ws = st.workspaces[0]          # of the first of the two compute streams
    ws = ws[(-ws.data_ptr()) % 256:]
    stats = _cabi.rdf_stats(ws, st.compute_stream.cuda_stream)
    hbm_peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm_peak = json.load(fh).get("hbm_gbs")
    except Exception:
        pass
    hbm_src = "MEASURED_PEAKS.json" if hbm_peak else "fallback 6650 GB/s"
    hbm_peak = hbm_peak or 6650.0
    alg_bytes = 12.0 * n1 * frames_all            # g1 == g2: read once
    # what the pair kernel actually evaluated in its last launch (one batch of
    # fps frames): the sort + box tests skip group pairs that provably hold no
    # pair in range, so evaluated << algorithmic
    frames_last = max(1, stats["tile_items"] // (-(-n1 // 128)))
    frames_per_launch = frames_last
    # DRAM traffic of the pair kernel per launch, from the committed ncu
    # capture (bytes per frame x frames of one launch); null when the workload
    # is not the one the capture was taken on
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1h_traffic.json")) as fh:
            tr = json.load(fh)
        if args.n_atoms == N_ATOMS:
            traffic = tr["dram_bytes_per_frame"] * frames_per_launch
    except Exception:
        pass
    eval_pairs_per_frame = stats["tile_pairs"] / float(frames_last)
    eval_frac = eval_pairs_per_frame / (float(n1) * n2)
    eval_rate = eval_pairs_per_frame * frames_all / kernel_s
    roofline = {
        "bound": "fp32", "kernel": "pair_prune_kernel<ORTHO>",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak, "traffic": traffic,
        "traffic_source": "profiles/r1h_traffic.json: dram__bytes_read+write of "
                          "one ncu --set full capture, per frame x frames per "
                          "launch",
        "peak_source": "FFMA loop measured live (rdfk_measure_fp32_peak); "
                       "MEASURED_PEAKS.json has no FP32 figure",
        "flop_per_pair": FLOP_PER_PAIR,
        "launch_ms": 1e3 * kernel_s / args.steps / world /
                     max(1.0, fps / float(frames_per_launch)),
        "frames_per_launch": frames_per_launch,
        "evaluated": {
            "pairs_last_launch": stats["tile_pairs"],
            "fraction_of_pair_matrix": eval_frac,
            "pairs_per_s": eval_rate,
            "tflops_at_8_flop_per_evaluated_pair": eval_rate * 8.0 / 1e12,
            "frac_of_peak": eval_rate * 8.0 / 1e12 / peak,
        },
        "fp64_rechecked_pairs_last_launch": stats["slow_pairs"],
        "hbm": {"algorithmic_GBps": alg_bytes / kernel_s / 1e9,
                "peak_GBps": hbm_peak, "peak_source": hbm_src,
                "frac": alg_bytes / kernel_s / 1e9 / hbm_peak},
        "note": "achieved = ALGORITHMIC flops (20 x N1 x N2 per frame, "
                "SURVEY 8d) / kernel time: exact pruning evaluates only "
                "`evaluated.fraction_of_pair_matrix` of the pair matrix "
                "(3 sub + 3 mul + 2 add = 8 flop per evaluated pair), so frac "
                "exceeds 1; the kernel is issue/latency bound, not FMA bound "
                "(profiles/r1_notes.md)",
    }

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle
        cpu = cpu_baseline_dict(args.n_atoms, c_oracle.max_threads(), 6.0)

    line = {
        "metric": "InterRDF pair-distances/sec", "value": value,
        "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64", "data": "synthetic",
        "frames_per_s": n_frames * args.steps / t_dev,
        "config": {"workload": WORKLOAD, "n_atoms": args.n_atoms,
                   "nbins": NBINS, "range": list(RANGE),
                   "frames_per_step_per_gpu": fps,
                   "l2": "inputs per step (%.0f MB/GPU) exceed the 126 MB L2"
                         % (fps * args.n_atoms * 12 / 1e6),
                   "parallelism": "frames sharded over %d rank(s), one int64 "
                                  "all-reduce per run()" % world},
        "e2e": {"value": e2e_value, "unit": "pairs/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * t_e2e / args.steps,
                "frames_per_s": n_frames * args.steps / t_e2e},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "count_sum": int(count_dev.sum()),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
