"""Where does run() spend host time on the small configurations?  Trace
points of the engine (engine._trace) averaged over repeated run()s of
BASELINE configs[0] and [2], resident and end to end."""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pmda_b200 import engine, synthetic  # noqa: E402
from pmda_b200.rdf import InterRDF, InterRDF_s  # noqa: E402


def main():
    which = [int(a) for a in sys.argv[1:]] or [1, 3]
    for cfg in which:
        nf = {1: 100, 3: 400}[cfg]
        u, kw = synthetic.sharded_config(cfg, nf)
        if cfg == 3:
            make = lambda: InterRDF_s(u, kw["ags"], nbins=kw["nbins"],
                                      range=kw["range"])
        else:
            make = lambda: InterRDF(kw["g1"], kw["g2"], nbins=kw["nbins"],
                                    range=kw["range"])
        for cache in (64 << 30, 0):
            engine.Config.cache_bytes = cache
            engine.clear_cache()
            for _ in range(5):
                make().run()
            torch.cuda.synchronize()
            reps = 50
            t0 = time.perf_counter()
            for _ in range(reps):
                r = make().run()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            print("cfg%d cache=%s: make().run() %.3f ms, kernels %.3f ms, "
                  "%d frames/s" % (cfg, bool(cache), dt * 1e3,
                                   1e3 * float(r.timing.compute.sum()),
                                   nf / dt), flush=True)
            acc = collections.OrderedDict()
            for _ in range(reps):
                engine._trace = tr = []
                t_a = time.perf_counter()
                a = make()
                tr.append(("analysis object", time.perf_counter()))
                a.run()
                engine._trace = None
                prev = t_a
                for label, t in tr:
                    acc[label] = acc.get(label, 0.0) + (t - prev)
                    prev = t
            for label, v in acc.items():
                print("    %-48s %7.1f us" % (label, 1e6 * v / reps))
            print("    %-48s %7.1f us" % ("(sum)", 1e6 * sum(acc.values()) / reps),
                  flush=True)


if __name__ == "__main__":
    main()
