"""A/B of PMDA_B200_COMPUTE_STREAMS (1 or 2): resident and end-to-end
frames/s of BASELINE configs 2, 4, 5 with several batches per run()."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmda_b200 import synthetic, engine
from pmda_b200.rdf import InterRDF
FRAMES = {2: 256, 4: 96, 5: 24}
out = []
for cfg in (2, 4, 5):
    u, kw = synthetic.CONFIGS[cfg](n_frames=FRAMES[cfg])
    make = lambda: InterRDF(kw["g1"], kw["g2"], nbins=kw["nbins"], range=kw["range"])
    res = []
    for cache in (64 << 30, 0):
        engine.Config.cache_bytes = cache
        engine.clear_cache(); make().run()
        best, kern = 1e9, 0
        for _ in range(4):
            r = make(); torch.cuda.synchronize(); t0 = time.perf_counter(); r.run(); torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if dt < best: best, kern = dt, float(r.timing.compute.sum())
        res.append((FRAMES[cfg] / best, FRAMES[cfg] / kern))
    out.append("cfg%d resident %.0f (timing %.0f) e2e %.0f" % (cfg, res[0][0], res[0][1], res[1][0]))
    del u, kw, r
print("STREAMS", os.environ.get("PMDA_B200_COMPUTE_STREAMS", "2"), " | ".join(out), flush=True)
