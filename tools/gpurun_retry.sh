#!/bin/bash
# usage: gpurun_retry.sh <timeout> <command...>   -- retries while the pod answers "transient" (no slot / busy)
T=$1; shift
for i in $(seq 1 12); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|status=busy"; then
    echo "[retry $i] no slot; sleeping"; sleep 150; continue
  fi
  echo "$OUT" | tail -40
  exit 0
done
echo "gave up"
