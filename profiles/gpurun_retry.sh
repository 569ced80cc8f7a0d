#!/bin/bash
# retries a gpurun call while the pod answers busy/transient (nothing is charged for those)
# usage: profiles/gpurun_retry.sh <timeout> <command...>
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_try.log 2>&1
  rc=$?
  if grep -q "status=ok\|status=fail\|status=timeout\|status=error" /tmp/gpurun_try.log; then cat /tmp/gpurun_try.log | tail -40; exit 0; fi
  if [ $rc -eq 2 ]; then cat /tmp/gpurun_try.log | tail; exit 2; fi
  tail -2 /tmp/gpurun_try.log
  sleep 90
done
exit 3
