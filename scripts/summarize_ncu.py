"""Summarise gpurun_out/prof.ncu-rep + launches.csv into profiles/ (run here, no GPU needed).
    python scripts/summarize_ncu.py <tag>      e.g.  r01_v3
Writes profiles/<tag>_ncu_full_summary.json, profiles/<tag>_launches.csv, profiles/<tag>_launch_shares.txt and
refreshes profiles/ncu_render_kernel_summary.json (read by bench.py for roofline.traffic)."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "latest"
rep = os.path.join(ROOT, "gpurun_out", "prof.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic"]
summ = {h: {"value": v, "unit": u} for h, u, v in zip(hdr, units, vals) if h in keep}


def to_bytes(d):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[d["unit"]]
    return float(d["value"]) * scale


dram = to_bytes(summ["dram__bytes_read.sum"]) + to_bytes(summ["dram__bytes_write.sum"])
summ["dram_bytes_per_launch"] = dram
summ["command"] = ("ncu --set full --clock-control none --import-source on -k regex:render_transient_path -s 1 -c 1 "
                   "python scripts/render_once.py --reps 2   (C2: 512x512x1024 bins, 256 spp, depth 8)")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(summ, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.json"), "w"), indent=1)
json.dump({"dram_bytes_per_launch": dram, "kernel": summ["Kernel Name"]["value"], "source": f"profiles/{tag}_ncu_full_summary.json",
           "duration_ms_under_ncu": float(summ["gpu__time_duration.sum"]["value"])},
          open(os.path.join(ROOT, "profiles", "ncu_render_kernel_summary.json"), "w"), indent=1)
lc = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(lc):
    shutil.copy(lc, os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
    r = list(csv.reader(open(lc)))
    i = [k for k, x in enumerate(r) if x and x[0] == "ID"][0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for x in r[i + 1:]:
        agg[x[4][:90]][0] += 1
        agg[x[4][:90]][1] += float(x[-1])
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launch_shares.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 1 --no-cpu-baseline\n")
        f.write("# (cold-cache, serialised launches: compare SHARES, not absolutes; includes the e2e leg's band kernels)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1] / 1e6:10.3f} ms  {v[0]:4d} launches  {100 * v[1] / tot:6.2f} %  {k}\n")
    print(open(os.path.join(ROOT, "profiles", f"{tag}_launch_shares.txt")).read())
print(json.dumps({k: v for k, v in summ.items() if k != "command"}, indent=0)[:1500])
