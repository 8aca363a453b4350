#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> '<command>' [--gpus N]
# Retries a gpurun call while the pod answers "transient" / busy (exit 3); stops on anything else.
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -40 /tmp/gpurun_last.log
exit $rc
