#!/bin/bash
# GPU job (one B200): parity tests of the new passes (one pytest process per group: a faulting kernel poisons its process only),
# compute-sanitizer on the dynamic-resolution tests
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_reblur.py tests/test_gpu_relax.py -m gpu -q -k "performance_mode or relax_per_pass or relax_settings or relax_sequence or optional or anti_firefly or one_signal_per_pass" --durations=4 > $O/r2_job11_tests.log 2>&1; tail -12 $O/r2_job11_tests.log | cut -c1-300
for d in REBLUR_DIFFUSE_SPECULAR RELAX_DIFFUSE_SPECULAR SIGMA_SHADOW; do
  timeout 300 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_aux.py -m gpu -q -x -k "dynamic_resolution and $d" > $O/r2_job11_dynres_$d.log 2>&1
  grep -E "Invalid|at 0x|by thread|in .*Kernel|passed|failed|ERROR SUMMARY|Address" $O/r2_job11_dynres_$d.log | head -14 | cut -c1-330
done
