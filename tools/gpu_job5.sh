#!/bin/bash
# GPU job: quick parity subset after kernel changes, A/B of the tap-batch variants, ncu of the spatial + TS kernels
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_gpu_reblur.py tests/test_gpu_strips.py tests/test_gpu_baseline_configs.py -m gpu -q -k "not config3 and not config4" --durations=5 > $O/r2_job5_tests.log 2>&1; tail -12 $O/r2_job5_tests.log
for v in "" tb1 tb2 tb4m3 tb8m3 tb2m5; do
  lib=raytracingdenoiser_b200/libnrd_b200${v:+_$v}.so
  NRD_B200_LIB=$PWD/$lib python bench.py --steps 12 --warmup 4 --no-cpu-baseline > $O/r2_ab_${v:-base}.json 2> $O/r2_ab_${v:-base}.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r2_ab_${v:-base}.json"))
    print("${v:-base}", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["per_pass_ms"].items()})
except Exception as e: print("${v:-base}", "failed", e)
PY
done
ncu --set full --clock-control none --import-source on -k regex:"ReblurSpatialKernel|TemporalStabilization" -s 20 -c 4 -o $O/r2_spatial_b -f python bench.py --steps 2 --warmup 4 --no-cpu-baseline > $O/r2_ncu1.log 2>&1; tail -2 $O/r2_ncu1.log
ls -la $O | grep r2_spatial_b
