#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3: nothing charged)
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
