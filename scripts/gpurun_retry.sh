#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> <gpus> <command...>   -- retries while the pod answers "busy" (exit code 3)
to=$1; shift; g=$1; shift
for attempt in $(seq 1 30); do
  if [ "$g" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $to -- "$@"; else /usr/local/graft/bin/gpurun --gpus $g --timeout $to -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy, attempt $attempt"; sleep 75
done
exit 3
