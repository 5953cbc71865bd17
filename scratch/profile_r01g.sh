#!/bin/bash
# bench (no profiler) -> ncu launch list -> ncu --set full of one launch of the main kernels; outputs under gpurun_out/
# (a capture of every launch with source is ~190 MB, over gpurun's 64 MiB return limit: the regex keeps it at ~16 kernels)
set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err || exit 1
tail -c 300 gpurun_out/bench_r01g.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01g.csv \
    python bench.py --steps 1 --warmup 1 --sims-per-step 592 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_g.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"sy2sb|sb2st|syrk|trmm|hrp_gram|hrp_tree_kernel<true>|hrp_tree_kernel<\(bool\)1>|eigvals|mpfit|markowitz|philox|reconstruct|backtransform" \
    -s 14 -c 14 -o gpurun_out/prof_r01g -f \
    python bench.py --steps 1 --warmup 1 --sims-per-step 592 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_g.log 2>&1
ls -la gpurun_out/
