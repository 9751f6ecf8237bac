#!/bin/bash
# ncu evidence of the round: launch list of the bench command + one full capture per workload (C2, C3, C4i).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/bench_under_ncu.log 2>&1
for wl in ${PROFILE_WORKLOADS:-cornell nlos vol}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_transient_path -s 1 -c 1 \
      -f -o gpurun_out/prof_$wl python scripts/time_kernel.py --scene $wl --steps 1 > gpurun_out/ncu_full_$wl.log 2>&1; echo "ncu $wl rc=$?"
done
ls -la gpurun_out/*.ncu-rep
