#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
for pass in 1 2; do for so in variants/*.so; do
  MITRANSIENT_B200_LIB=$PWD/$so python scripts/time_kernel.py 2>&1 | tail -1
  MITRANSIENT_B200_LIB=$PWD/$so python scripts/time_kernel.py --spp 2048 --world 8 2>&1 | tail -1
done; done
for so in variants/*.so; do MITRANSIENT_B200_LIB=$PWD/$so python scripts/time_kernel.py --scene nlos 2>&1 | tail -1; done
