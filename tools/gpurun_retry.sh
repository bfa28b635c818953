#!/bin/bash
# gpurun with retries while the pod answers "transient" / busy (nothing charged); usage: gpurun_retry.sh <timeout> [--gpus N] -- '<command>'
T=$1; shift
for i in $(seq 1 14); do
    OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1)
    echo "$OUT" | tail -150
    if echo "$OUT" | grep -q "status=ok\|status=fail\|status=error\|status=timeout"; then exit 0; fi
    if ! echo "$OUT" | grep -q "transient\|busy\|rc=3\|no box"; then exit 0; fi
    echo "[retry $i] waiting 150 s"; sleep 150
done
