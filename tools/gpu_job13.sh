#!/bin/bash
# Final GPU job of the round (one B200): RELAX 4K per-pass parity (the RELAX kernels changed), N=1 bench lines (product + reference arm),
# ncu launch list of the denoiser kernels and ncu --set full of one steady-state REBLUR frame
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 170 python bench.py --steps 20 --warmup 5 > $O/r2_final_bench_n1.json 2> $O/r2_final_bench_n1.err; tail -c 1500 $O/r2_final_bench_n1.json; tail -2 $O/r2_final_bench_n1.err
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_final_bench_ref.json 2> $O/r2_final_bench_ref.err; tail -c 400 $O/r2_final_bench_ref.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"Reblur|Clear" -c 200 --csv --log-file $O/r2_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/r2_launches.log 2>&1; tail -1 $O/r2_launches.log | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"Reblur" -s 98 -c 7 -o $O/r2_reblur_final -f python bench.py --steps 12 --warmup 4 --no-cpu-baseline > $O/r2_ncu13.log 2>&1; tail -1 $O/r2_ncu13.log
NRD_B200_GPU_TEST_BUDGET_S=100000 timeout 500 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -k "config4" > $O/r2_job13_config4.log 2>&1; tail -5 $O/r2_job13_config4.log | cut -c1-300
ls -la $O | grep -E "final|launches|ncu-rep" | cut -c1-120
