#!/bin/bash
# Round-end GPU session: GPU tests, random GPU-vs-oracle sweep, smoke, bench (N = 1), ncu evidence.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 1200 python tests/gpu_fuzz.py ${FUZZ_N:-300} ${FUZZ_SEED:-0} > gpurun_out/gpu_fuzz.log 2>&1; echo "fuzz rc=$?" | tee -a gpurun_out/gpu_fuzz.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
bash scripts/gpu_profile.sh > gpurun_out/profile.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; grep -c "^OK" gpurun_out/gpu_fuzz.log; grep "^FAIL\|^#" gpurun_out/gpu_fuzz.log | head -20; tail -2 gpurun_out/smoke.log; cut -c1-600 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
