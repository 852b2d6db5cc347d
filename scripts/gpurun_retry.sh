#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <command string>   -- retries while the pod answers "busy" (exit 3)
t=$1; shift
for attempt in $(seq 1 20); do
    /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 90
done
exit 3
