#!/bin/bash
# development aid: retry a gpurun call while the pod answers "busy" (exit code 3); usage: tools/gpurun_retry.sh TIMEOUT 'command'
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
