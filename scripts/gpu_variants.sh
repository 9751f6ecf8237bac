#!/bin/bash
# A/B timing of variants/*.so against the default library on the scenes in $VARIANT_SCENES (default: cornell nlos)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; P=gpurun_out/progress.log; : > $P
run() { echo "== $*" >> $P; timeout 120 "$@" >> $P 2>&1; echo "rc=$?" >> $P; }
for sc in ${VARIANT_SCENES:-cornell nlos}; do run python scripts/time_kernel.py --scene $sc; done
for so in variants/*.so; do [ -f $so ] || continue; for sc in ${VARIANT_SCENES:-cornell nlos}; do MITRANSIENT_B200_LIB=$PWD/$so run python scripts/time_kernel.py --scene $sc; done; done
grep -v "^$\|^rc=0\|^==" $P
