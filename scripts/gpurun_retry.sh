#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3, nothing charged).  usage: gpurun_retry.sh <timeout_s> <tag> <command...>
T=$1; TAG=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > gpurun_out/${TAG}_call.txt 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $i tries"; exit $rc; fi
  sleep 60
done
echo "gave up"; exit 3
