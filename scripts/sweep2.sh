#!/bin/bash
# A/B timing of variants/*.so at 256 spp (C2) and at 2048 spp on 1/8 of the pixels (the per-GPU share of an 8-GPU run)
cd "${GRAFT_REPO_ROOT:-.}"
for pass in 1 2; do for so in variants/*.so; do
  MITRANSIENT_B200_LIB=$PWD/$so python scripts/time_kernel.py 2>&1 | tail -1
  MITRANSIENT_B200_LIB=$PWD/$so python scripts/time_kernel.py --height 64 --spp 2048 2>&1 | tail -1
done; done
