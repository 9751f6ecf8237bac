cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for sc in cornell nlos vol; do python scripts/time_kernel.py --scene $sc; done > gpurun_out/time_base.log 2>&1
MITRANSIENT_B200_LIB=$PWD/variants/vol640.so python scripts/time_kernel.py --scene vol > gpurun_out/time_variants.log 2>&1
cat gpurun_out/time_base.log gpurun_out/time_variants.log
bash scripts/gpu_profile.sh
