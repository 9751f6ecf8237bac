#!/bin/bash
# A/B timing of prebuilt kernel variants (variants/*.so), two passes to expose run-to-run noise
cd "${GRAFT_REPO_ROOT:-.}"
for pass in 1 2; do for so in variants/*.so; do MITRANSIENT_B200_LIB=$PWD/$so python scripts/time_kernel.py 2>&1 | tail -1; done; done
for so in variants/*.so; do echo $so; MITRANSIENT_B200_LIB=$PWD/$so python scripts/render_nlos.py 2>&1 | tail -1 | cut -c1-60; done
