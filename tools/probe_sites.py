"""InterRDF_s throughput on the BASELINE config 3 shape."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pmda_b200 import synthetic, engine
from pmda_b200.rdf import InterRDF_s
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 400
u, a = synthetic.config3(n_frames=nf)
for cache in (48 << 30, 0):
    engine.Config.cache_bytes = cache
    engine.clear_cache()
    for rep in range(3):
        r = InterRDF_s(u, a["ags"], nbins=a["nbins"], range=a["range"])
        torch.cuda.synchronize(); t0 = time.time(); r.run(); torch.cuda.synchronize(); dt = time.time() - t0
        print("cache=%d rep%d: wall %.4fs  %.0f frames/s  kernel %.4fs  io %.4fs  count.sum=%d" % (
            cache > 0, rep, dt, nf / dt, r.timing.compute.sum(), r.timing.io.sum(),
            int(sum(c.sum() for c in r.count))), flush=True)
