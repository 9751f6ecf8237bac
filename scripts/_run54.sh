cd "${GRAFT_REPO_ROOT:-.}"
for so in default variants/spare12.so variants/spare20.so; do
  for sc in "cornell 256" "cornell 64" "cornell 1024" "mesh 255" "sphere3 257"; do set -- $sc
    if [ $so = default ]; then unset MITRANSIENT_B200_LIB; else export MITRANSIENT_B200_LIB=$PWD/$so; fi
    echo -n "$so $1 $2: "; timeout 120 python scripts/time_kernel.py --scene $1 --spp $2 --steps 5 2>&1 | tail -1 | awk '{print $2, $3}'
  done
done
