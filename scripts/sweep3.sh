#!/bin/bash
# per-GPU share of the 8-GPU weak-scaling job on one GPU: 1/8 of the tiles at 2048 spp, for variants/*.so
cd "${GRAFT_REPO_ROOT:-.}"
for pass in 1 2; do for so in variants/*.so; do
  MITRANSIENT_B200_LIB=$PWD/$so python scripts/time_kernel.py --spp 2048 --world 8 2>&1 | tail -1
done; done
python scripts/time_kernel.py --spp 1024 --world 4 | tail -1
python scripts/time_kernel.py --spp 256 --world 1 | tail -1
