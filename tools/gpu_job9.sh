#!/bin/bash
# GPU job (one B200), perf first: A/B of the spatial-kernel build variants, timings of every chain, ncu of the REBLUR and RELAX chains;
# then the small-size parity tests of the new passes.  Every step has its own timeout (the host CPU of the box may be slow).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for v in "" mb3 mb3b2 mb2b8; do
  lib=raytracingdenoiser_b200/libnrd_b200${v:+_$v}.so
  NRD_B200_LIB=$PWD/$lib timeout 170 python bench.py --steps 12 --warmup 4 --no-cpu-baseline > $O/r2_ab9_${v:-base}.json 2> $O/r2_ab9_${v:-base}.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$O/r2_ab9_${v:-base}.json") if l.startswith("{")][-1]
    print("${v:-base}", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["per_pass_ms"].items()})
except Exception as e: print("${v:-base}", "failed", e)
PY
done
timeout 240 python tools/time_chains.py --frames 10 --warmup 8 > $O/r2_all_chains_4k.txt 2> $O/r2_all_chains_4k.err; cut -c1-420 $O/r2_all_chains_4k.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"Reblur" -s 98 -c 7 -o $O/r2_reblur_steady2 -f python bench.py --steps 12 --warmup 4 --no-cpu-baseline > $O/r2_ncu9a.log 2>&1; tail -1 $O/r2_ncu9a.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"Relax" -s 99 -c 11 -o $O/r2_relax_steady -f python tools/time_chains.py --only RELAX_DIFFUSE_SPECULAR --frames 4 --warmup 8 > $O/r2_ncu9b.log 2>&1; tail -1 $O/r2_ncu9b.log
timeout 420 python -m pytest tests/test_gpu_aux.py tests/test_gpu_relax.py tests/test_gpu_strips.py -m gpu -q -x -k "aux or one_signal or strips_bit_identical" --durations=5 > $O/r2_job9_tests.log 2>&1; tail -12 $O/r2_job9_tests.log
ls -la $O | grep -E "ncu-rep"
