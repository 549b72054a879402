#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit code 3):  tools/gpurun_retry.sh [--gpus N] --timeout S -- 'cmd'
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
