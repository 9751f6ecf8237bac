#!/bin/bash
# Build the current tree with extra nvcc flags into variants/<name>.so:   scripts/build_variant.sh name [-DTR_...=...]
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p variants
TR_NVCC_EXTRA="$*" python -m mitransient_b200.build --force > /dev/null
cp mitransient_b200/lib/libmitransient_b200.so ${VARIANT_DIR:-variants}/$name.so
grep -A3 "ILb1ELb1ELb1ELi0ELb0" mitransient_b200/lib/build.log | grep -E "spill|Used" | tr '\n' ' '; echo " <- $name"
