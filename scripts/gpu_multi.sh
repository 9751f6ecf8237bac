#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the 2-GPU pytest + bench.py under torchrun at N ranks.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$N.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpus_n$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_multi_n$N.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 ) > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/pytest_multi_n$N.log; cat gpurun_out/bench_n$N.json; tail -8 gpurun_out/bench_n$N.err
