#!/bin/bash
# usage: scripts/grunN.sh <n_gpus> <timeout_s> <logname> <command...>  -- retries while the pod answers busy (exit 3 / transient)
N=$1; shift; T=$1; shift; LOG=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@" > "gpurun_out/$LOG.gpurun.log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "gpurun_out/$LOG.gpurun.log"; then exit $rc; fi
  sleep 60
done
exit 3
