#!/bin/bash
# GPU job: TMA probe, the complete GPU suite, the N = 1 bench line, A/B variants, ncu captures of the spatial / TA / HF kernels
cd "$(dirname "$0")/.."
O=gpurun_out
tools/tma_probe > $O/r2_tma_probe.txt 2>&1; tail -2 $O/r2_tma_probe.txt
python -m pytest tests -m gpu -q --durations=12 > $O/r2_gputest_full.log 2>&1; tail -25 $O/r2_gputest_full.log
python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 1800 $O/r2_bench_n1.json; tail -3 $O/r2_bench_n1.err
for v in "" tacall b5 notma; do
  lib=raytracingdenoiser_b200/libnrd_b200${v:+_$v}.so
  if [ "$v" = notma ]; then lib=raytracingdenoiser_b200/libnrd_b200.so; export NRD_B200_NO_TMA=1; fi
  NRD_B200_LIB=$PWD/$lib python bench.py --steps 12 --warmup 4 --no-cpu-baseline > $O/r2_ab_${v:-base}.json 2> $O/r2_ab_${v:-base}.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r2_ab_${v:-base}.json"))
    print("${v:-base}", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["per_pass_ms"].items()})
except Exception as e: print("${v:-base}", "failed", e)
PY
done
unset NRD_B200_NO_TMA
ncu --set full --clock-control none --import-source on -k regex:ReblurSpatialKernel -s 9 -c 3 -o $O/r2_spatial -f python bench.py --steps 2 --warmup 4 --no-cpu-baseline > $O/r2_ncu1.log 2>&1; tail -2 $O/r2_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:"TemporalAccumulation|HistoryFix|TemporalStabilization" -s 9 -c 3 -o $O/r2_temporal -f python bench.py --steps 2 --warmup 4 --no-cpu-baseline > $O/r2_ncu2.log 2>&1; tail -2 $O/r2_ncu2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_ncu3.log 2>&1; tail -2 $O/r2_ncu3.log
ls -la $O | grep r2_
