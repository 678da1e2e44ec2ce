#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3: nothing charged).
#   tools/gpurun_retry.sh [--gpus N] [--timeout S] -- '<command>'
for attempt in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[gpurun_retry] no slot (attempt $attempt), sleeping 90 s" >&2
    sleep 90
done
exit 3
