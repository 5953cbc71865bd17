#!/bin/bash
# A/B of the thread count of one-CTA-per-simulation kernels (timing only; restores the committed build afterwards)
cp mcos_b200/libmcosb200.so /tmp/lib_good.so
for spec in "denoise CORR_ROWS 64" "denoise CORR_ROWS 16" "hrp HD_ROWS 64" "hrp HD_ROWS 16"; do
  set -- $spec
  rm -f mcos_b200/csrc/$1.o
  make -C mcos_b200/csrc -j8 EXTRA=-D$2=$3 > /dev/null 2>&1
  echo "== $2=$3"
  timeout 200 python scratch/kt.py 8192 2>&1 | grep -E "value|corr|hrp_dist"
  rm -f mcos_b200/csrc/$1.o
done
cp /tmp/lib_good.so mcos_b200/libmcosb200.so
