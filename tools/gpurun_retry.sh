#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <script> [gpus]   -- retries while the pod answers busy/transient
T=$1; S=$2; G=${3:-1}
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  if [ "$G" = 1 ]; then out=$(gpurun --timeout $T -- "bash $S" 2>&1); else out=$(gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  echo "$out" | tail -40
  if echo "$out" | grep -q 'status=transient\|status=busy\|rc=None'; then sleep 120; continue; fi
  break
done
