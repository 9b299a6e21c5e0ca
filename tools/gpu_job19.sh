#!/bin/bash
# Round 2, job 19 (the last ~100 GPU-seconds): the checkerboard gates again after the checkerboard code of the temporal-accumulation
# kernels became a template parameter (default kernels compile as before), plus the default per-pass gates in the same process.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 60 python -m pytest tests/test_gpu_checkerboard.py tests/test_gpu_reblur.py tests/test_gpu_relax.py -m gpu -q -k "checkerboard or test_reblur_per_pass_parity or test_relax_per_pass_parity" > $O/r2_job19.log 2>&1
tail -3 $O/r2_job19.log
timeout 25 python tools/time_chains.py --frames 8 --warmup 4 --only RELAX_DIFFUSE_SPECULAR > $O/r2_job19_chains.txt 2>&1
tail -1 $O/r2_job19_chains.txt | cut -c1-330
