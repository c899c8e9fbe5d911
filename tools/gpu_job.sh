#!/usr/bin/env bash
# developer aid: run a script on the GPU box, retrying while the pod has no free slot
# usage: tools/gpu_job.sh <timeout_s> <script-on-the-box> [gpus]
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = 1 ]; then out=$(gpurun --timeout "$T" -- "bash $S" 2>&1); else out=$(gpurun --gpus "$G" --timeout "$T" -- "bash $S" 2>&1); fi
  if echo "$out" | grep -q "status=transient\|status=busy\|no box"; then sleep 90; continue; fi
  echo "$out" | tail -150
  exit 0
done
echo "gave up: pod busy"
