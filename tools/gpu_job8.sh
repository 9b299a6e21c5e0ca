#!/bin/bash
# GPU job (one B200): parity of the re-written spatial filters / optional inputs / RELAX anti-firefly, A/B of the spatial-kernel build variants, timings of the other chains,
# ncu of the RELAX chain
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_reblur.py tests/test_gpu_relax.py tests/test_gpu_baseline_configs.py -m gpu -q -x -k "not config3 and not config4 and not sequence and not golden" --durations=8 > $O/r2_job8_tests.log 2>&1; tail -15 $O/r2_job8_tests.log
for v in "" mb3 mb3b2 mb2b8 mb4b2; do
  lib=raytracingdenoiser_b200/libnrd_b200${v:+_$v}.so
  NRD_B200_LIB=$PWD/$lib python bench.py --steps 12 --warmup 4 --no-cpu-baseline > $O/r2_ab8_${v:-base}.json 2> $O/r2_ab8_${v:-base}.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$O/r2_ab8_${v:-base}.json") if l.startswith("{")][-1]
    print("${v:-base}", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["per_pass_ms"].items()})
except Exception as e: print("${v:-base}", "failed", e)
PY
done
python tools/time_chains.py --frames 12 --warmup 8 > $O/r2_all_chains_4k.txt 2> $O/r2_all_chains_4k.err; cat $O/r2_all_chains_4k.txt | cut -c1-700
ncu --set full --clock-control none --import-source on -k regex:"Relax" -s 88 -c 11 -o $O/r2_relax_steady -f python tools/time_chains.py --only RELAX_DIFFUSE_SPECULAR --frames 4 --warmup 8 > $O/r2_ncu8.log 2>&1; tail -2 $O/r2_ncu8.log
ls -la $O | grep -E "ncu-rep|ab8"
