"""Hydrogen-bond search throughput on one GPU: water
boxes of 3,000 (BASELINE config 1's system with its hydrogens) and 33,000
molecules, device-resident and end to end through HydrogenBondAnalysis.run,
plus the oracle's CPU time per frame on the small box.  Documentation numbers
for DESIGN.md / profiles; not the contract bench (bench.py)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pmda_b200 import synthetic, engine
from pmda_b200.hbond_analysis import HydrogenBondAnalysis
from pmda_b200.universe import Universe

CASES = [("water3k", 3000, 44.78, 256), ("water33k", 33000, 99.6, 32)]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if c[0] in sys.argv[1:]]


def timed(make, reps=3):
    best, r = 1e9, None
    for _ in range(reps):
        r = make()
        torch.cuda.synchronize(); t0 = time.perf_counter(); r.run(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, r


for name, n_mol, L, nf in CASES:
    rng = np.random.default_rng(3)
    dims = [L, L, L, 90, 90, 90]
    pos1, names = synthetic.water_box(1, n_mol, dims, rng)
    # frames = jittered copies of one configuration (cheap to generate)
    pos = pos1 + rng.normal(0, 0.02, (nf,) + pos1.shape[1:]).astype(np.float32)
    u = Universe(pos.astype(np.float32), dims, names=names)
    o = 3 * np.arange(n_mol)
    don, hyd = np.repeat(o, 2), (o[:, None] + np.array([1, 2])).reshape(-1)
    make = lambda: HydrogenBondAnalysis(u, acceptors_sel=o, pairs=(don, hyd))
    engine.Config.cache_bytes = 64 << 30
    engine.clear_cache()
    make().run()
    t_dev, r = timed(make)
    engine.Config.cache_bytes = 0
    engine.clear_cache()
    make().run()
    t_e2e, r = timed(make)
    screened = float(len(don)) * len(o)
    res = {"molecules": n_mol, "frames": nf, "hbonds_per_frame": len(r.hbonds) / nf,
           "screened_pairs_per_frame": screened,
           "device_resident_frames_per_s": nf / t_dev,
           "e2e_frames_per_s": nf / t_e2e,
           "kernel_frames_per_s": nf / float(r.timing.compute.sum()),
           "kernel_screened_pairs_per_s": nf * screened / float(r.timing.compute.sum())}
    if n_mol <= 3000:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle import c_oracle, numpy_oracle
        ts = u.trajectory[0]
        t0 = time.perf_counter()
        rows = numpy_oracle.hbonds_single_frame(
            0, ts.positions, don, hyd, o, ts.dimensions,
            distance_fn=c_oracle.distance_array)
        res["cpu_oracle_frames_per_s"] = 1.0 / (time.perf_counter() - t0)
        res["cpu_oracle_threads"] = c_oracle.max_threads()
        got = r.hbonds[r.hbonds[:, 0] == 0]
        assert np.array_equal(got[:, :5], rows[:, :5])
    print(name, json.dumps(res), flush=True)
    del u, r
