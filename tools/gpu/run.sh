#!/bin/bash
# usage: tools/gpu/run.sh <timeout_s> <script> [gpus]   -- retries while the pod answers "transient"
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 12); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $S" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  echo "$OUT" | tail -12
  if echo "$OUT" | grep -q "status=transient"; then echo "[retry $i]"; sleep 90; continue; fi
  break
done
