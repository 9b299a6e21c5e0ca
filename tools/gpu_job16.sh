#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_aux.py tests/test_gpu_reblur.py -m gpu -q -k "wider_formats or base_color or reblur_per_pass_parity" > $O/r2_job16.log 2>&1; tail -8 $O/r2_job16.log | cut -c1-400
