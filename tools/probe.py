"""Quick GPU probe: throughput of the tile kernel on BASELINE config shapes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pmda_b200 import synthetic, engine, _cabi
from pmda_b200.rdf import InterRDF

def run(cfg, n_frames, n_atoms=None, reps=3):
    kw = {} if n_atoms is None else {"n_atoms": n_atoms}
    u, a = synthetic.CONFIGS[cfg](n_frames=n_frames, **kw)
    r = InterRDF(a["g1"], a["g2"], nbins=a["nbins"], range=a["range"])
    for rep in range(reps):
        t0 = time.time(); r.run(); dt = time.time() - t0
        pairs = r.n_frames * len(a["g1"]) * len(a["g2"])
        ker = r.timing.compute.sum()
        print("cfg%d F=%d N=%dx%d rep%d: wall %.4fs (%.3e pairs/s)  kernel %.4fs (%.3e pairs/s, %.1f frames/s)  count.sum=%d" % (
            cfg, n_frames, len(a["g1"]), len(a["g2"]), rep, dt, pairs/dt, ker, pairs/ker, r.n_frames/ker, r.count.sum()), flush=True)
    st = engine._state(torch.device("cuda", torch.cuda.current_device()))
    print("  stats:", _cabi.rdf_stats(st.workspaces[0][(-st.workspaces[0].data_ptr()) % 256:], st.compute_stream.cuda_stream))

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), _cabi.device_info(0))
    print("fp32 peak TFLOP/s:", _cabi.measure_fp32_peak())
    run(1, 100)
    run(2, 4)
    run(2, 16)
    run(4, 2, n_atoms=100_000)
    run(5, 8)
