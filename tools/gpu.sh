#!/bin/bash
# gpurun with retries on "no slot right now" (exit 3): tools/gpu.sh [--gpus N] [--timeout S] -- 'command'
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
