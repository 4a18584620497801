"""Frames/s and pair-distances/s of the five BASELINE.json configurations on
one GPU: device-resident (frame cache hit) and end to end from host memory.
Not the contract bench (bench.py); documentation numbers for DESIGN.md."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pmda_b200 import synthetic, engine
from pmda_b200.rdf import InterRDF, InterRDF_s

FRAMES = {1: 100, 2: 128, 3: 400, 4: 48, 5: 24}


def timed(make, reps=4):
    best = 1e9
    for _ in range(reps):
        r = make()
        torch.cuda.synchronize(); t0 = time.perf_counter(); r.run(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, r


out = {}
for cfg in (1, 2, 3, 4, 5):
    nf = FRAMES[cfg]
    u, kw = synthetic.CONFIGS[cfg](n_frames=nf)
    if cfg == 3:
        make = lambda: InterRDF_s(u, kw["ags"], nbins=kw["nbins"], range=kw["range"])
        pairs = sum(len(a) * len(b) for a, b in kw["ags"])
    else:
        make = lambda: InterRDF(kw["g1"], kw["g2"], nbins=kw["nbins"], range=kw["range"])
        pairs = len(kw["g1"]) * len(kw["g2"])
    engine.Config.cache_bytes = 64 << 30
    engine.clear_cache()
    make().run()
    t_dev, r = timed(make)
    engine.Config.cache_bytes = 0
    engine.clear_cache()
    make().run()
    t_e2e, r = timed(make)
    out[cfg] = {"frames": nf, "pairs_per_frame": pairs,
                "device_resident_frames_per_s": nf / t_dev, "e2e_frames_per_s": nf / t_e2e,
                "device_resident_pairs_per_s": nf * pairs / t_dev,
                "e2e_pairs_per_s": nf * pairs / t_e2e,
                "kernel_frames_per_s": nf / float(r.timing.compute.sum())}
    print("cfg%d" % cfg, json.dumps(out[cfg]), flush=True)
    del u, kw, r
