#!/bin/bash
# GPU job: the complete GPU suite (BASELINE-size parity gates included) and the N = 1 bench line
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q --durations=12 > $O/r2_gputest_full.log 2>&1; tail -25 $O/r2_gputest_full.log
python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 1500 $O/r2_bench_n1.json
