#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>' [extra gpurun args...]
# Retries while the pod answers "busy" (exit code 3: nothing charged), every 120 s, up to 40 times.
T=$1; shift; CMD=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
