#!/bin/bash
# usage: [GPUS=2] tools/gpurun_retry.sh <timeout_s> '<command>'  -- retries while the pod answers "transient" / busy (exit 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$T" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 150; continue; fi
  cat /tmp/gpurun_last.log | tail -80
  exit $rc
done
echo "gave up after 40 tries"; exit 3
