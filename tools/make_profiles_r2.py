#!/usr/bin/env python
"""Regenerates the round-2 evidence under profiles/ from the captures gpurun brought back in gpurun_out/ (run here, on the CPU box):
ncu summaries (metric table + stall reasons + hottest SASS lines + instruction mix), profiles/r2_traffic.json (read by bench.py for
roofline.traffic / issue-active), launch lists, bench lines."""
import csv, json, os, shutil, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(REPO, "gpurun_out"); P = os.path.join(REPO, "profiles")
py = sys.executable


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(d, k):
    v, u = d[k]
    f = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}
    return f * scale.get(u, 1.0)


def summary(rep, dst, header):
    parts = [header, ""]
    for tool, args in (("ncu_summary.py", []), ("ncu_hot.py", ["25"]), ("ncu_mix.py", [])):
        parts.append("## tools/%s" % tool)
        parts.append(subprocess.run([py, os.path.join(REPO, "tools", tool), rep] + args, capture_output=True, text=True).stdout)
    open(dst, "w").write("\n".join(parts))


traffic = {}
caps = [("r2_prof_gd_project.ncu-rep", "r2_ncu_gd_project.txt", "k_gd_project",
         "ncu --set full --clock-control none --import-source on of k_gd_project<0,2> (unrolled walks, joints 5/6 merged), 1e6 S-GD samples, `python bench.py --steps 2 --warmup 3 --no-extra --no-cpu`",
         "1e6 S-GD samples per launch; algorithmic bytes 96e6 in + 96e6 raw joints out + 4e6 state = 196e6"),
        ("r2_prof_collide.ncu-rep", "r2_ncu_collide.txt", "k_collide",
         "ncu --set full of k_collide<0> in the GD workload (31.5 k configurations per launch; link frames computed in the warp, heavy routine at the end of the same launch)",
         "31.5k configurations per launch (GD workload); algorithmic bytes = 31.5e3 x (96 B joints in + 1 B mask out) = 3.1e6; the rest is the hierarchy / triangle working set entering L2 once per (cold, profiler-serialised) launch"),
        ("r2_prof_collide_pcs.ncu-rep", "r2_ncu_collide_pcs.txt", "k_collide_pcs",
         "ncu --set full of k_collide<0> in the PCS workload (one launch per 1e6-sample batch: ~24 k configurations)", "24k configurations per launch (PCS workload, one launch per 1e6-sample batch); see k_collide for the DRAM note")]
for rep, dst, key, header, note in caps:
    rp = os.path.join(G, rep)
    if not os.path.exists(rp):
        print("missing", rep); continue
    summary(rp, os.path.join(P, dst), "# " + header)
    d = raw(rp)
    t = {"dram_bytes_read": num(d, "dram__bytes_read.sum"), "dram_bytes_write": num(d, "dram__bytes_write.sum")}
    t["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"]
    t["duration_s"] = num(d, "gpu__time_duration.sum")
    t["issue_active_pct"] = float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0])
    t["lanes_per_instruction"] = float(d["smsp__thread_inst_executed_per_inst_executed.ratio"][0])
    t["fp64_pipe_active_pct"] = float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0])
    t["warp_instructions"] = float(d["smsp__inst_executed.sum"][0].replace(",", ""))
    t["registers_per_thread"] = int(d["launch__registers_per_thread"][0])
    t["capture"] = rep; t["note"] = note
    traffic[key] = t
    print(key, json.dumps(t)[:300])
if traffic:
    json.dump(traffic, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)
for f in ("r2_launches_gd.csv", "r2_launches_pcs.csv", "r2_launches_rbs.csv"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
for src, dst in (("bench_default.json", "r2_bench_default.json"), ("bench_default_ref.json", "r2_bench_default_ref.json"), ("r2_bench_gd_n8.json", "r2_bench_gd_n8.json"),
                 ("e2e_sweep2.txt", "r2_e2e_pipeline_sweep.txt"), ("h2d_multi.txt", "r2_h2d_concurrent_ranks.txt"), ("multi_probe.txt", "r2_multi_device_context.txt"),
                 ("park.txt", "r2_gd_parking_rejected.txt"), ("k3b.txt", "r2_k3_defer_thresholds.txt")):
    if os.path.exists(os.path.join(G, src)):
        lines = open(os.path.join(G, src)).read().strip().split("\n")
        if src.endswith(".json"):
            lines = [l for l in lines if l.startswith("{")][-1:]
        open(os.path.join(P, dst), "w").write("\n".join(lines) + "\n")
print("done")
