#!/bin/bash
# Quick ncu A/B of prebuilt kernel variants (variants/*.so): instruction counts, issue utilisation, local-memory traffic.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,gpu__time_duration.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
for so in variants/*.so; do
  n=$(basename $so .so)
  MITRANSIENT_B200_LIB=$PWD/$so ncu --metrics $M --clock-control none -k regex:render_transient_path -s 1 -c 1 --csv --log-file gpurun_out/ab_$n.csv python scripts/time_kernel.py --steps 1 > /dev/null 2>&1
  echo "== $n"; tail -n +2 gpurun_out/ab_$n.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"'
done
if [ -n "$1" ]; then
  MITRANSIENT_B200_LIB=$PWD/variants/$1.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_transient_path -s 1 -c 1 -f -o gpurun_out/prof_$1 python scripts/time_kernel.py --steps 1 > gpurun_out/ncu_full_$1.log 2>&1; echo "ncu full rc=$?"
fi
