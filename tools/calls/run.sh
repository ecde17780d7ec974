#!/bin/bash
# usage: tools/calls/run.sh <script> <timeout_s>  -- retries while the pod answers busy/transient
for i in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun --timeout $2 -- "bash $1" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3"; then sleep 120; continue; fi
  break
done
