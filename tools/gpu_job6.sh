#!/bin/bash
# GPU job: parity subset after the HistoryFix compaction, A/B (strip build on one GPU, TA occupancy variants), steady-state ncu of the temporal kernels
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_gpu_reblur.py tests/test_gpu_strips.py tests/test_gpu_baseline_configs.py -m gpu -q -k "not config3 and not config4" --durations=5 > $O/r2_job6_tests.log 2>&1; tail -12 $O/r2_job6_tests.log
for v in "" strip ta6 ta7 ta8; do
  lib=raytracingdenoiser_b200/libnrd_b200${v:+_$v}.so
  unset NRD_B200_FORCE_STRIP_KERNELS
  if [ "$v" = strip ]; then lib=raytracingdenoiser_b200/libnrd_b200.so; export NRD_B200_FORCE_STRIP_KERNELS=1; fi
  NRD_B200_LIB=$PWD/$lib python bench.py --steps 12 --warmup 4 --no-cpu-baseline > $O/r2_ab_${v:-base}.json 2> $O/r2_ab_${v:-base}.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$O/r2_ab_${v:-base}.json") if l.startswith("{")][-1]
    print("${v:-base}", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["per_pass_ms"].items()})
except Exception as e: print("${v:-base}", "failed", e)
PY
done
unset NRD_B200_FORCE_STRIP_KERNELS
ncu --set full --clock-control none --import-source on -k regex:"TemporalAccumulation|HistoryFix|TemporalStabilization" -s 42 -c 3 -o $O/r2_temporal_steady -f python bench.py --steps 12 --warmup 4 --no-cpu-baseline > $O/r2_ncu2.log 2>&1; tail -2 $O/r2_ncu2.log
ls -la $O | grep r2_temporal_steady
