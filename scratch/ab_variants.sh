#!/bin/bash
# A/B timing of library variants built with scratch/build_variant.sh: prints value and the kernels named in $KEYS for each
KEYS=${KEYS:-"eigvals q2back hrp_tree reconstruct"}
for name in base "$@"; do
  lib=$PWD/mcos_b200/libmcosb200_$name.so
  [ "$name" = base ] && lib=$PWD/mcos_b200/libmcosb200.so
  MCOSB_LIB_PATH=$lib python bench.py --no-configs --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > gpurun_out/ab_$name.json 2>/dev/null
  python - "$name" $KEYS <<PY
import json, sys
d = json.loads(open("gpurun_out/ab_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"], 1), {k: d["kernels"][k]["ms"] for k in sys.argv[2:] if k in d["kernels"]})
PY
done
