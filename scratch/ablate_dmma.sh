#!/bin/bash
# timing experiment: SYRK / TRMM with parts of the k-step removed (results are wrong; only the kernel times matter)
cp mcos_b200/libmcosb200.so /tmp/lib_good.so
for v in 4 7; do
  rm -f mcos_b200/csrc/moments.o mcos_b200/csrc/sample.o
  make -C mcos_b200/csrc -j8 EXTRA=-DDM_ABLATE=$v > /dev/null 2>&1
  echo "== DM_ABLATE=$v"
  timeout 200 python scratch/kt.py 8192 2>&1 | grep -E "syrk|trmm"
done
rm -f mcos_b200/csrc/moments.o mcos_b200/csrc/sample.o
cp /tmp/lib_good.so mcos_b200/libmcosb200.so
