"""End-to-end runs at (or towards) the trajectory lengths BASELINE.json names,
frames streamed from host memory (device frame cache off):

    cfg2: 1000 frames x 100k atoms (as specified, 1.2 GB)
    cfg3: 2000 frames x 50k atoms  (as specified, 1.2 GB; InterRDF_s)
    cfg4:  500 of the 5000 frames x 250k atoms (1.5 GB per GPU-sized piece)
    cfg5:  100 of the 10000 frames x 1M atoms (1.2 GB)

One JSON line per configuration (profiles/r2_full_length.json)."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pmda_b200 import engine, synthetic  # noqa: E402
from pmda_b200.rdf import InterRDF, InterRDF_s  # noqa: E402

FRAMES = {2: 1000, 3: 2000, 4: 500, 5: 100}
which = [int(a) for a in sys.argv[1:]] or [2, 3, 4, 5]
engine.Config.cache_bytes = 0
for cfg in which:
    nf = FRAMES[cfg]
    t0 = time.perf_counter()
    u, kw = synthetic.CONFIGS[cfg](n_frames=nf)
    t_gen = time.perf_counter() - t0
    if cfg == 3:
        make = lambda: InterRDF_s(u, kw["ags"], nbins=kw["nbins"], range=kw["range"])
        pairs = sum(len(a) * len(b) for a, b in kw["ags"])
    else:
        make = lambda: InterRDF(kw["g1"], kw["g2"], nbins=kw["nbins"], range=kw["range"])
        pairs = len(kw["g1"]) * len(kw["g2"])
    engine.clear_cache()
    make().run()                               # warm-up (buffers, pinning)
    best, r = 1e9, None
    for _ in range(2):
        r = make()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r.run()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({
        "config": cfg, "frames": nf, "pairs_per_frame": pairs,
        "host_bytes": int(u.trajectory.positions.nbytes),
        "e2e_seconds": best, "e2e_frames_per_s": nf / best,
        "e2e_pairs_per_s": nf * pairs / best,
        "kernel_frames_per_s": nf / float(r.timing.compute.sum()),
        "generate_seconds": t_gen}), flush=True)
    del u, kw, r
