#!/bin/bash
# One ncu --set full capture of the render kernel for a time_kernel.py scene:  gpu_profile_one.sh <scene> [extra args]
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
wl=$1; shift
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_transient_path -s 1 -c 1 \
    -f -o /tmp/prof_$wl python scripts/time_kernel.py --scene $wl --steps 1 "$@" > gpurun_out/ncu_full_$wl.log 2>&1; echo "ncu $wl rc=$?"
ncu -i /tmp/prof_$wl.ncu-rep --page raw --csv > gpurun_out/prof_${wl}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$wl.ncu-rep --page source --csv > gpurun_out/prof_${wl}_source.csv 2>/dev/null
ls -la gpurun_out/prof_${wl}_*
