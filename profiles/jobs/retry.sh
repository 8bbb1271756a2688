#!/bin/bash
# retry.sh <timeout_s> <job script>: gpurun with back-off while the pod answers "busy" (exit 3)
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2"; rc=$?
  if [ $rc -ne 3 ]; then break; fi
  sleep 90
done
echo "final rc=$rc"
