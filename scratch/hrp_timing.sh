#!/bin/bash
cp mcos_b200/libmcosb200.so /tmp/lib_good.so
rm -f mcos_b200/csrc/hrp.o
make -C mcos_b200/csrc -j8 EXTRA=-DHRP_TIMING > /dev/null 2>&1
timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --sims-per-step 4096 2>&1 | grep "hrp_tree<" | head
rm -f mcos_b200/csrc/hrp.o
cp /tmp/lib_good.so mcos_b200/libmcosb200.so
