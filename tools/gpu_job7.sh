#!/bin/bash
# GPU job (one B200): whole GPU suite, N=1 bench line, ncu launch list, ncu --set full of one steady-state frame of the REBLUR chain
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q --durations=12 > $O/r2_job7_tests.log 2>&1; tail -25 $O/r2_job7_tests.log
python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 2500 $O/r2_bench_n1.json; tail -3 $O/r2_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err; tail -c 600 $O/r2_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2_launches.log 2>&1; tail -2 $O/r2_launches.log
ncu --set full --clock-control none --import-source on -k regex:"Reblur" -s 98 -c 7 -o $O/r2_reblur_steady -f python bench.py --steps 12 --warmup 4 --no-cpu-baseline > $O/r2_ncu7.log 2>&1; tail -2 $O/r2_ncu7.log
ls -la $O
