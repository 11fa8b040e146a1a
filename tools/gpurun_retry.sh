#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <script> : resubmits while the pod answers "busy" (exit code 3)
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
