"""Summarise an ncu --set full capture of the render kernel into profiles/ (run here, no GPU needed).
    python scripts/summarize_ncu.py <tag> [workload=C2] [report=gpurun_out/prof.ncu-rep] [command text]
e.g.  python scripts/summarize_ncu.py r02_v3 C3 gpurun_out/prof_C3.ncu-rep
Writes profiles/<tag>_ncu_<workload>_summary.json and refreshes profiles/ncu_<workload>_summary.json (read by bench.py for
roofline.traffic / roofline.issue); for C2 also profiles/<tag>_launches.csv + _launch_shares.txt from gpurun_out/launches.csv."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "latest"
wl = sys.argv[2] if len(sys.argv) > 2 else "C2"
rep = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "prof.ncu-rep")
cmd = sys.argv[4] if len(sys.argv) > 4 else None
if rep.endswith(".csv"):          # already converted on the GPU box (ncu -i ... --page raw --csv), see scripts/gpu_profile.sh
    out = open(rep).read()
else:
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
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed_op_branch.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "sm__cycles_active.avg", "smsp__cycles_active.avg"]
keep += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
summ = {h: {"value": v, "unit": u} for h, u, v in zip(hdr, units, vals) if h in keep}


def to_bytes(d):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[d["unit"]]
    return float(d["value"]) * scale


dram = to_bytes(summ["dram__bytes_read.sum"]) + to_bytes(summ["dram__bytes_write.sum"])
summ["dram_bytes_per_launch"] = dram
summ["workload"] = wl
summ["kernel_version"] = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE, text=True).stdout.strip()
summ["command"] = cmd or ("ncu --set full --clock-control none --import-source on -k regex:render_transient_path -s 1 -c 1 "
                          "python scripts/render_once.py --reps 2   (C2: 512x512x1024 bins, 256 spp, depth 8)")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(summ, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_{wl}_summary.json"), "w"), indent=1)
json.dump(summ, open(os.path.join(ROOT, "profiles", f"ncu_{wl}_summary.json"), "w"), indent=1)
lc = os.path.join(ROOT, "gpurun_out", "launches.csv")
if wl == "C2" and os.path.exists(lc):
    shutil.copy(lc, os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
    r = list(csv.reader(open(lc)))
    i = [k for k, x in enumerate(r) if x and x[0] == "ID"][0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for x in r[i + 1:]:
        agg[x[4][:90]][0] += 1
        agg[x[4][:90]][1] += float(x[-1])
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launch_shares.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra\n")
        f.write("# (cold-cache, serialised launches: compare SHARES, not absolutes; includes the e2e leg's band kernels)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1] / 1e6:10.3f} ms  {v[0]:4d} launches  {100 * v[1] / tot:6.2f} %  {k}\n")
    print(open(os.path.join(ROOT, "profiles", f"{tag}_launch_shares.txt")).read())
for k, v in summ.items():
    if isinstance(v, dict):
        print(f"{k:90s} {v['value']} {v['unit']}")
