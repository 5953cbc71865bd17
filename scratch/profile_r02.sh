#!/bin/bash
# bench (no profiler) -> ncu launch list -> ncu --set full of one launch of the main kernels; outputs under gpurun_out/ (<= 64 MiB)
# usage: profile_r02.sh a | b
set -x
export PYTHONPATH=$PWD
if [ "$1" = "a" ]; then
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err || exit 1
tail -c 300 gpurun_out/bench_r02c.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 1 --warmup 1 --sims-per-step 592 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/ncu_launch_r02.log 2>&1
ncu --set full --clock-control none \
    -k regex:"sy2sb|sb2st|syrk|trmm|hrp_gram|eigvals|mpfit|markowitz|philox|reconstruct|backtransform|q2back|corr_kernel|hrp_prim" \
    -s 14 -c 14 -o gpurun_out/prof_r02 -f \
    python bench.py --steps 1 --warmup 1 --sims-per-step 592 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/ncu_full_r02.log 2>&1
else
# config 5: the N = 1000 kernels (sy2sb2, nco_cluster, nco_alloc), and the dominant kernel with source
ncu --set full --clock-control none -k regex:"sy2sb2|nco_cluster|nco_alloc" -c 3 -o gpurun_out/prof_r02_n1000 -f \
    python scratch/config5_profile.py 296 > gpurun_out/ncu_full_r02_n1000.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sy2sb" -s 1 -c 1 -o gpurun_out/prof_r02_sy2sb -f \
    python bench.py --steps 1 --warmup 1 --sims-per-step 592 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/ncu_full_r02_sy2sb.log 2>&1
fi
ls -la gpurun_out/ | tail -8
