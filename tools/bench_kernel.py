"""Kernel-only and device-resident frames/s of selected BASELINE configs
(default 2 and 4) for the library RDFK_LIB points at: the A/B harness for
kernel experiments (`python -m pmda_b200.csrc.build --out=... -DRDFK_VAR=n`)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmda_b200 import synthetic, engine
from pmda_b200.rdf import InterRDF

FRAMES = {1: 100, 2: 128, 4: 48, 5: 24}
cfgs = [int(a) for a in sys.argv[1:]] or [2, 4]
engine.Config.cache_bytes = 64 << 30
if os.environ.get("RDFK_FLAGS"):   # rdfk_set_debug_flags (A/B of code paths)
    from pmda_b200 import _cabi
    _cabi.set_debug_flags(int(os.environ["RDFK_FLAGS"], 0))
out = []
for cfg in cfgs:
    u, kw = synthetic.CONFIGS[cfg](n_frames=FRAMES[cfg])
    make = lambda: InterRDF(kw["g1"], kw["g2"], nbins=kw["nbins"], range=kw["range"])
    engine.clear_cache()
    make().run()
    best, kern = 1e9, 1e9
    for _ in range(5):
        r = make()
        torch.cuda.synchronize(); t0 = time.perf_counter(); r.run(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
        kern = min(kern, float(r.timing.compute.sum()))
    out.append("cfg%d resident %.0f kernel %.0f" % (cfg, FRAMES[cfg] / best, FRAMES[cfg] / kern))
    del u, kw, r
print(os.environ.get("RDFK_LIB", "default"), os.environ.get("RDFK_FLAGS", ""),
      " | ".join(out), flush=True)
