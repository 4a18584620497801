"""Turn the raw outputs of the evidence run (gpurun_out/<tag>_*) into the
tracked files under profiles/:

    python tools/collect_evidence.py r2z

  profiles/r2_bench_n1.json          bench.py line (1 GPU), verbatim
  profiles/r2_configs.json           the per-config table (GPU: bench.py's
                                     `configs` + headline; CPU: cpu_configs.py;
                                     contacts / hydrogen bonds)
  profiles/r2_pair_prune_{ortho,tri}_ncu_full_summary.csv, r2_sites_ncu_full_summary.csv
  profiles/r2_pair_prune_{ortho,tri}_sass_hotspots.txt
  profiles/r2_traffic.json           DRAM bytes per frame + issue-slot use of
                                     the ortho capture (bench.py reads it)
  profiles/r2_launches_bench_small.csv
  profiles/r2_full_length.json       tools/bench_full_length.py, one line per config
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2z"


def last_json(path):
    with open(path) as fh:
        lines = [l for l in fh.read().splitlines() if l.strip().startswith("{")]
    return json.loads(lines[-1])


def json_lines(path):
    out = []
    if not os.path.exists(path):
        return out
    for l in open(path).read().splitlines():
        i = l.find("{")
        if i >= 0:
            try:
                d = json.loads(l[i:])
            except ValueError:
                continue
            if i > 0:
                d["case"] = l[:i].strip()
            out.append(d)
    return out


def run(cmd, dest):
    with open(dest, "w") as fh:
        subprocess.run(cmd, stdout=fh, check=True)


bench = last_json(os.path.join(OUT, tag + "_bench.json"))
with open(os.path.join(PROF, "r2_bench_n1.json"), "w") as fh:
    json.dump(bench, fh, indent=1)
table = {
    "source": "bench.py (GPU rows), tools/cpu_configs.py (CPU rows), "
              "tools/bench_contacts.py, tools/bench_hbonds.py; one gpurun call",
    "gpu": {"cfg2_headline": {
        "workload": bench["config"]["workload"],
        "resident_frames_per_s": bench["frames_per_s"],
        "e2e_frames_per_s": bench["e2e"]["frames_per_s"],
        "resident_pairs_per_s": bench["value"],
        "e2e_pairs_per_s": bench["e2e"]["value"],
        "evaluated_fraction":
            bench["roofline"]["evaluated"]["fraction_of_pair_matrix"]}},
    "cpu_baseline_cfg2": bench.get("cpu_baseline"),
    "cpu": json_lines(os.path.join(OUT, tag + "_cpu_configs.txt")),
    "contacts": json_lines(os.path.join(OUT, tag + "_contacts.txt")),
    "hbonds": json_lines(os.path.join(OUT, tag + "_hbonds.txt")),
}
table["gpu"].update(bench.get("configs") or {})
for n in (2, 4, 8):
    p = os.path.join(OUT, "%s_bench_n%d.json" % (tag, n))
    if os.path.exists(p):
        d = last_json(p)
        with open(os.path.join(PROF, "r2_bench_n%d.json" % n), "w") as fh:
            json.dump(d, fh, indent=1)
with open(os.path.join(PROF, "r2_configs.json"), "w") as fh:
    json.dump(table, fh, indent=1)

for kind in ("ortho", "tri", "sites"):
    rep = os.path.join(OUT, "%s_%s.ncu-rep" % (tag, kind))
    if not os.path.exists(rep):
        continue
    name = "r2_sites" if kind == "sites" else "r2_pair_prune_" + kind
    run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep],
        os.path.join(PROF, name + "_ncu_full_summary.csv"))
    src = "/tmp/%s_%s_src.csv" % (tag, kind)
    run(["ncu", "-i", rep, "--page", "source", "--csv"], src)
    run([sys.executable, os.path.join(ROOT, "tools", "ncu_sass_summary.py"), src],
        os.path.join(PROF, name + "_sass_hotspots.txt"))

# DRAM traffic + issue-slot use of the ortho capture
rep = os.path.join(PROF, "r2_pair_prune_ortho_ncu_full_summary.csv")
if os.path.exists(rep):
    m = {r[0]: r for r in csv.reader(open(rep))}

    def val(k):
        return float(m[k][2])

    def mb(k):
        v, unit = float(m[k][2]), m[k][1].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
    frames = int(open(os.path.join(OUT, tag + "_ortho_frames.txt")).read())
    traffic = {
        "capture": "ncu --set full --clock-control none, one "
                   "pair_prune_kernel<ORTHO> launch of `python "
                   "tools/bench_kernel.py 2` (100k atoms, %d frames)" % frames,
        "frames_in_launch": frames,
        "dram_bytes_read": mb("dram__bytes_read.sum"),
        "dram_bytes_write": mb("dram__bytes_write.sum"),
        "dram_bytes_per_frame":
            (mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")) / frames,
        "algorithmic_bytes_per_frame": 12.0 * 100000,
        "ipc": val("sm__inst_executed.avg.per_cycle_elapsed"),
        "issue_slot_frac":
            val("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
        "shared_pipe_frac": val(
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"
            ".pct_of_peak_sustained_elapsed") / 100.0,
        "shared_wavefronts_per_frame":
            val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / frames,
        "warp_instructions_per_frame": val("smsp__inst_executed.sum") / frames,
        "kernel_ms": val("gpu__time_duration.sum") *
            {"ms": 1.0, "us": 1e-3, "s": 1e3, "msecond": 1.0, "usecond": 1e-3,
             "second": 1e3}.get(m["gpu__time_duration.sum"][1], 1.0),
    }
    with open(os.path.join(PROF, "r2_traffic.json"), "w") as fh:
        json.dump(traffic, fh, indent=1)
for name, dest in ((tag + "_bench_reference.json", "r2_bench_reference_n1.json"),
                   (tag + "_bench_reference_n8.json", "r2_bench_reference_n8.json")):
    p = os.path.join(OUT, name)
    if os.path.exists(p):
        try:
            with open(os.path.join(PROF, dest), "w") as fh:
                json.dump(last_json(p), fh, indent=1)
        except Exception:
            pass
src = os.path.join(OUT, tag + "_full_length.txt")
if os.path.exists(src):
    with open(src) as a, open(os.path.join(PROF, "r2_full_length.json"), "w") as b:
        b.write(a.read())
src = os.path.join(OUT, tag + "_launches_bench_small.csv")
if os.path.exists(src):
    with open(src) as a, open(os.path.join(PROF, "r2_launches_bench_small.csv"), "w") as b:
        b.write(a.read())
print("profiles/ updated from", tag)
