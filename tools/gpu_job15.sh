#!/bin/bash
# GPU job (one B200): the tests added after job 12 (wider user formats, split screen / REFERENCE regression), quick
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 240 python -m pytest tests/test_gpu_aux.py -m gpu -q -k "wider_formats or split_screen or reference_denoiser" > $O/r2_job15_aux.log 2>&1; tail -12 $O/r2_job15_aux.log | cut -c1-300
