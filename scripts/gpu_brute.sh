#!/bin/bash
# where does the BVH overtake the lock-step loop?  (TR_BRUTE_MAX overrides the 64-primitive threshold)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; P=gpurun_out/progress.log; : > $P
run() { echo "== $TR_BRUTE_MAX $*" >> $P; timeout 120 "$@" >> $P 2>&1; echo "rc=$?" >> $P; }
for lim in 64 8; do export TR_BRUTE_MAX=$lim; run python scripts/time_kernel.py --scene mesh --spp 256 --steps 3; done
for lim in 64 128; do export TR_BRUTE_MAX=$lim; run python scripts/time_kernel.py --scene sphere0 --steps 3; run python scripts/time_kernel.py --scene sphere1 --steps 3; done
grep -v "^$\|^rc=0" $P
