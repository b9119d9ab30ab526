#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit code 3 / status=transient): nothing is charged for those.
# Usage: tools/gpurun_retry.sh <tries> <gpurun args...>
tries=$1; shift
for i in $(seq 1 $tries); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|status=busy\|retry in a few minutes"; then
    echo "[retry $i] pod busy"; sleep 150; continue
  fi
  echo "$out"; exit $rc
done
echo "gave up after $tries tries"; exit 3
