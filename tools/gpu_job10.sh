#!/bin/bash
# GPU job (one B200): small-size parity tests of the passes added / changed since job 9 (performance mode, dynamic resolution, RELAX A-trous rewrite,
# optional inputs), then timings of every chain.  Every step has its own timeout.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_aux.py tests/test_gpu_reblur.py tests/test_gpu_relax.py -m gpu -q -k "dynamic_resolution or performance_mode or relax_per_pass or relax_settings or relax_sequence or optional or anti_firefly or one_signal_per_pass" --durations=6 > $O/r2_job10_tests.log 2>&1; tail -30 $O/r2_job10_tests.log | cut -c1-400
timeout 240 python tools/time_chains.py --frames 10 --warmup 8 > $O/r2_all_chains_4k_b.txt 2> $O/r2_all_chains_4k_b.err; cut -c1-330 $O/r2_all_chains_4k_b.txt; tail -2 $O/r2_all_chains_4k_b.err
