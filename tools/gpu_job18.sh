#!/bin/bash
# Round 2, job 18 (last GPU minutes of the round): checkerboarded inputs (REBLUR + RELAX) against the oracle, then a regression
cd "$(dirname "$0")/.."
# pass over the default per-pass gates and the chain timings.  Every step has its own timeout; no -x so that all results are seen.
O=gpurun_out
mkdir -p $O
timeout 100 python -m pytest tests/test_gpu_checkerboard.py -m gpu -q -k "reblur" --durations=3 > $O/r2_job18_reblur.log 2>&1
tail -5 $O/r2_job18_reblur.log
timeout 100 python -m pytest tests/test_gpu_checkerboard.py -m gpu -q -k "relax" --durations=3 > $O/r2_job18_relax.log 2>&1
tail -5 $O/r2_job18_relax.log
timeout 80 python -m pytest tests/test_gpu_reblur.py tests/test_gpu_relax.py -m gpu -q -k "test_reblur_per_pass_parity or test_relax_per_pass_parity or test_relax_sequence" > $O/r2_job18_regress.log 2>&1
tail -3 $O/r2_job18_regress.log
timeout 70 python tools/time_chains.py --frames 8 --warmup 4 --only REBLUR_DIFFUSE_SPECULAR,RELAX_DIFFUSE_SPECULAR > $O/r2_job18_chains.txt 2>&1
tail -4 $O/r2_job18_chains.txt
