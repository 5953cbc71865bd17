#!/bin/bash
# Final round-2 evidence: bench (no profiler) -> ncu launch list of the same command shape -> ncu --set full of one launch of
# the main kernels (the largest pass / panel launches of the band reduction included).  Outputs under gpurun_out/ (<= 64 MiB).
set -x
export PYTHONPATH=$PWD
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err || exit 1
tail -c 300 gpurun_out/bench_r02f.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02f.csv \
    python bench.py --steps 1 --warmup 1 --sims-per-step 592 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/ncu_launch_r02f.log 2>&1
# --set full: (a) the first pass / panel launches of the second step (largest trailing matrix), (b) one launch of each other kernel
ncu --set full --clock-control none --import-source on -k regex:"s3_pass|s3_panel" -s 126 -c 4 -o gpurun_out/prof_r02f_s3 -f \
    python bench.py --steps 1 --warmup 1 --sims-per-step 592 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/ncu_full_r02f_s3.log 2>&1
ncu --set full --clock-control none \
    -k regex:"sb2st|syrk|trmm|hrp_gram|eigvals|mpfit|markowitz|philox|reconstruct|backtransform|q2back|corr_kernel|hrp_prim|hrp_dist|hrp_tree|inviter" \
    -s 20 -c 18 -o gpurun_out/prof_r02f -f \
    python bench.py --steps 1 --warmup 1 --sims-per-step 592 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/ncu_full_r02f.log 2>&1
# gpurun_out/ travels back only below 64 MiB: summarise the big report here and keep the CSV, not the report
python tools/ncu_summaries.py kernels gpurun_out/prof_r02f.ncu-rep > gpurun_out/kernels_r02f.md
python tools/ncu_summaries.py kernels gpurun_out/prof_r02f_s3.ncu-rep > gpurun_out/kernels_r02f_s3.md
ncu -i gpurun_out/prof_r02f.ncu-rep --page raw --csv > gpurun_out/kernels_r02f_raw.csv
rm -f gpurun_out/prof_r02f.ncu-rep
ls -la gpurun_out/ | tail -12
