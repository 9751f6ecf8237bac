#!/bin/bash
# BVH path: timings on the icosphere scenes (256 spp) for the default library and variants/*.so, lanes per instruction from ncu
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; P=gpurun_out/progress.log; : > $P
run() { echo "== $*" >> $P; timeout 120 "$@" >> $P 2>&1; echo "rc=$?" >> $P; }
for sc in sphere2 sphere3 mesh vol nlos; do run python scripts/time_kernel.py --scene $sc --steps 3; done
for so in variants/*.so; do [ -f $so ] || continue; for sc in sphere2 sphere3; do MITRANSIENT_B200_LIB=$PWD/$so run python scripts/time_kernel.py --scene $sc --steps 3; done; done
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
for sc in sphere3 mesh; do timeout 300 ncu --metrics $M --clock-control none -k regex:render_transient_path -s 1 -c 1 --csv --log-file gpurun_out/lanes_$sc.csv python scripts/time_kernel.py --scene $sc --steps 1 > /dev/null 2>&1; echo "-- ncu $sc" >> $P; tail -n 4 gpurun_out/lanes_$sc.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' >> $P; done
run python -m pytest tests -q -m gpu -x
grep -v "^$" $P
