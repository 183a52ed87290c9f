#!/bin/bash
# usage: gpurun_retry.sh <out-file> <gpurun args...>  -- retries while the pod answers "transient" (nothing charged)
out=$1; shift
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  if ! grep -q "status=transient" "$out"; then exit 0; fi
  sleep 150
done
