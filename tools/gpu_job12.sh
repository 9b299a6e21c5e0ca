#!/bin/bash
# GPU job (one B200): parity of the re-written RELAX passes and of the remaining new passes, one pytest process per group; RELAX timing
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_relax.py -m gpu -q --durations=3 > $O/r2_job12_relax.log 2>&1; tail -6 $O/r2_job12_relax.log | cut -c1-300
timeout 200 python -m pytest tests/test_gpu_aux.py -m gpu -q -k "dynamic_resolution and REBLUR" > $O/r2_job12_dynres.log 2>&1; tail -4 $O/r2_job12_dynres.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_strips.py tests/test_gpu_baseline_configs.py -m gpu -q -k "strips_bit_identical or strip_build_per_pass or agree_within" > $O/r2_job12_strips.log 2>&1; tail -4 $O/r2_job12_strips.log | cut -c1-300
timeout 200 python tools/time_chains.py --only RELAX_DIFFUSE_SPECULAR --frames 10 --warmup 8 > $O/r2_relax_chain.txt 2>&1; tail -1 $O/r2_relax_chain.txt | cut -c1-400
