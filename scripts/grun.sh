#!/bin/bash
# usage: scripts/grun.sh <timeout_s> <logname> <command...>  -- retries while the pod answers "busy" (exit 3)
T=$1; shift; LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "gpurun_out/$LOG.gpurun.log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
