#!/bin/bash
# Round 2, first GPU session: GPU tests of the merged specular branches + baseline timings of the three workloads.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
for sc in cornell nlos vol; do python scripts/time_kernel.py --scene $sc; done > gpurun_out/time_base.log 2>&1
for so in variants/*.so; do
  for sc in cornell nlos vol; do MITRANSIENT_B200_LIB=$PWD/$so python scripts/time_kernel.py --scene $sc; done
done > gpurun_out/time_variants.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/time_base.log gpurun_out/time_variants.log
