#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  Usage: scripts/gpurun_retry.sh [gpurun args] -- 'command'
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry $i] pod busy, sleeping 90 s" >&2
  sleep 90
done
exit 3
