#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- '<command>'   -- retries while gpurun answers "no box / slot free" (exit 3)
T=$1; shift
for attempt in $(seq 1 30); do
    /usr/local/graft/bin/gpurun --timeout "$T" "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[retry] attempt $attempt answered busy, sleeping 90 s"
    sleep 90
done
exit 3
