"""CPU side of the per-config table: the oracle's
restatement of the reference's per-frame path on the host cores, on a bounded
row sample of one frame of each BASELINE.json configuration, extrapolated to
frames/s.  `cells`: through the oracle's cell list (what MDAnalysis'
capped_distance dispatch does at these sizes; orthorhombic only); `matrix`:
the full distance_array pair matrix.  TEST-INFRASTRUCTURE timing only."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from pmda_b200 import synthetic
from oracle import c_oracle

T = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)


def sample(pos, i1, i2, dims, nbins, rng, rows, threads, cells):
    g1, g2 = pos[i1][:rows], pos[i2]
    t0 = time.perf_counter()
    c_oracle.rdf_single_frame(g1, g2, dims, nbins, rng, nthreads=threads, cells=cells)
    dt = time.perf_counter() - t0
    return dt * len(i1) / rows            # seconds per full frame


for cfg, rows_m, rows_c in ((1, 3000, 3000), (2, 600, 6000), (4, 200, 0), (5, 100, 1000)):
    u, kw = synthetic.CONFIGS[cfg](n_frames=1)
    ts = u.trajectory[0]
    i1, i2 = kw["g1"].indices, kw["g2"].indices
    out = {"config": cfg, "threads": T, "pairs_per_frame": float(len(i1)) * len(i2)}
    sample(ts.positions, i1, i2, ts.dimensions, kw["nbins"], kw["range"], 50, T, False)
    s = sample(ts.positions, i1, i2, ts.dimensions, kw["nbins"], kw["range"], rows_m, T, False)
    out["matrix_frames_per_s"] = 1.0 / s
    if cfg == 1:
        s4 = sample(ts.positions, i1, i2, ts.dimensions, kw["nbins"], kw["range"], rows_m, 4, False)
        out["matrix_frames_per_s_4_threads"] = 1.0 / s4
    if rows_c:
        s = sample(ts.positions, i1, i2, ts.dimensions, kw["nbins"], kw["range"], rows_c, T, True)
        out["cells_frames_per_s"] = 1.0 / s
    print(json.dumps(out), flush=True)
    del u, kw

# config 3: InterRDF_s, 64 site pairs (1 x 256), the C restatement of the
# per-pair one-hot histogram (the reference's Python loop is far slower)
u, kw = synthetic.config3(n_frames=4)
t0 = time.perf_counter()
for f in range(4):
    ts = u.trajectory[f]
    pairs = [(ts.positions[a.indices], ts.positions[b.indices]) for a, b in kw["ags"]]
    c_oracle.rdf_s_single_frame(pairs, ts.dimensions, kw["nbins"], kw["range"])
print(json.dumps({"config": 3, "threads": 1, "c_restatement_frames_per_s": 4 / (time.perf_counter() - t0)}), flush=True)
