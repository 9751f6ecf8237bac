#!/bin/bash
# pixel-affine shared-memory histograms (in flight <= slots x spp) vs global vector reductions into the film
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; P=gpurun_out/progress.log; : > $P
t() { for h in ${HISTS:-shared global}; do echo "== $1 spp $2 $h" >> $P; timeout 120 python scripts/time_kernel.py --scene $1 --spp $2 --steps 5 --hist $h 2>&1 | tail -1 >> $P; done; }
for sc in ${SCENES:-vol:8 vol:16 vol:32 vol:64 nlos:384}; do t ${sc%%:*} ${sc##*:}; done
grep -v "^$" $P
