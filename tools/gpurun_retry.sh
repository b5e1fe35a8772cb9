#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> '<command>'   -- retries while the pod answers "busy / no box" (exit 3), nothing is charged for those
T=$1; G=$2; CMD=$3
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD"; fi
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  echo "[retry $i] rc=$rc, sleeping 120 s"; sleep 120
done
exit 3
