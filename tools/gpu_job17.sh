#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 260 python -m pytest tests/test_gpu_sigma.py tests/test_gpu_properties.py tests/test_gpu_reblur.py tests/test_api.py -m gpu -q -x -k "not per_pass_parity and not performance_mode_per_pass" --durations=4 > $O/r2_job17.log 2>&1; tail -10 $O/r2_job17.log | cut -c1-300
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
