#!/bin/bash
# GPU job on N GPUs of one box: cross-process bit-identity of the strips (tests/multi_gpu_check.py) and the bench line at N (and N/2)
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-2}
: > $O/r2_multi_check_final_n$N.log
CFGS=("REBLUR_DIFFUSE_SPECULAR 3840 2160 3" "RELAX_DIFFUSE_SPECULAR 1920 1080 3")
if [ "$N" -ge 8 ]; then CFGS=("REBLUR_DIFFUSE_SPECULAR 3840 2160 3"); fi
for cfg in "${CFGS[@]}"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py $cfg >> $O/r2_multi_check_final_n$N.log 2>&1
  echo "rc=$? $cfg" >> $O/r2_multi_check_final_n$N.log
done
grep -E "^rc=|strips_vs_full_frame" $O/r2_multi_check_final_n$N.log | cut -c1-300
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_bench_final_n$N.json 2> $O/r2_bench_final_n$N.err; tail -c 1200 $O/r2_bench_final_n$N.json; tail -3 $O/r2_bench_final_n$N.err
if [ "$N" -ge 4 ]; then
  H=$((N/2))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $H --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $H --steps 20 --warmup 5 > $O/r2_bench_n$H.json 2> $O/r2_bench_n$H.err; tail -c 600 $O/r2_bench_n$H.json
fi
