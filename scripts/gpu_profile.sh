#!/bin/bash
# ncu evidence of the round: launch list of the bench command + one full capture per workload (C2, C3, C4i).  The reports
# are converted to CSV on the box (raw page; source page for C2) and deleted: gpurun only brings back 64 MiB.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/bench_under_ncu.log 2>&1
for wl in ${PROFILE_WORKLOADS:-cornell nlos vol}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_transient_path -s 1 -c 1 \
      -f -o /tmp/prof_$wl python scripts/time_kernel.py --scene $wl --steps 1 > gpurun_out/ncu_full_$wl.log 2>&1; echo "ncu $wl rc=$?"
  ncu -i /tmp/prof_$wl.ncu-rep --page raw --csv > gpurun_out/prof_${wl}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$wl.ncu-rep --page source --csv > gpurun_out/prof_${wl}_source.csv 2>/dev/null
done
ls -la gpurun_out/
