#!/bin/bash
# GPU job: strips + REBLUR parity tests, issue-rate microbenchmark, tap-unroll A/B, ncu captures of the spatial + TA kernels
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q -k "not config3 and not config4 and not config5" > $O/r2_job2_tests.log 2>&1; tail -15 $O/r2_job2_tests.log
tools/ubench_issue > $O/r2_ubench_issue.txt 2>&1; cat $O/r2_ubench_issue.txt
for v in u1 "" u4 u8; do
  lib=raytracingdenoiser_b200/libnrd_b200${v:+_$v}.so
  NRD_B200_LIB=$PWD/$lib python bench.py --steps 12 --warmup 4 --no-cpu-baseline > $O/r2_ab_${v:-u2}.json 2> $O/r2_ab_${v:-u2}.err
  python - <<PY
import json
d=json.load(open("$O/r2_ab_${v:-u2}.json"))
print("${v:-u2}", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["per_pass_ms"].items()})
PY
done
ncu --set full --clock-control none --import-source on -k regex:ReblurSpatialKernel -s 9 -c 3 -o $O/r2_spatial -f python bench.py --steps 2 --warmup 4 --no-cpu-baseline > $O/r2_ncu1.log 2>&1; tail -2 $O/r2_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:TemporalAccumulation -s 3 -c 1 -o $O/r2_ta -f python bench.py --steps 2 --warmup 4 --no-cpu-baseline > $O/r2_ncu2.log 2>&1; tail -2 $O/r2_ncu2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_launches_mid.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_ncu3.log 2>&1; tail -2 $O/r2_ncu3.log
ls -la $O | grep r2_
