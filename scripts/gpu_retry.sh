#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers busy)
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@"
  rc=$?
  st=$(python -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json')).get('status'))" 2>/dev/null)
  if [ "$st" != "transient" ] && [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpu_retry] busy (attempt $i), sleeping 60 s"
  sleep 60
done
exit 3
