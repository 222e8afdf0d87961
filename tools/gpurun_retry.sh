#!/bin/bash
# Retries a gpurun call while the pod answers "busy / draining" (exit 3 or status=transient), at most N times.
# usage: tools/gpurun_retry.sh <timeout_s> <tries> <command string> [extra gpurun args...]
T=$1; N=$2; CMD=$3; shift 3
for i in $(seq 1 $N); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T "$@" -- "$CMD" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|no box or slot\|retry in a few minutes"; then
    echo "[retry $i] pod busy"; sleep 90; continue
  fi
  echo "$OUT" | tail -120
  exit 0
done
echo "gave up after $N tries"
