#!/bin/bash
# Quick GPU session: GPU tests + kernel timings of the three workloads (default lib and variants/*.so).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
for sc in cornell nlos vol; do timeout 120 python scripts/time_kernel.py --scene $sc; done > gpurun_out/time_base.log 2>&1
for so in variants/*.so; do
  [ -f "$so" ] || continue
  for sc in ${VARIANT_SCENES:-cornell}; do MITRANSIENT_B200_LIB=$PWD/$so timeout 120 python scripts/time_kernel.py --scene $sc; done
done > gpurun_out/time_variants.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/time_base.log gpurun_out/time_variants.log
