#!/bin/bash
# usage: tools/gpu_bg.sh <tag> <timeout_s> '<command>'   -- runs gpurun in the background, retrying while the pod is busy;
# log in /tmp/gpurun_<tag>.log, done marker /tmp/gpurun_<tag>.done
tag=$1; to=$2; cmd=$3
rm -f /tmp/gpurun_$tag.done
(
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to $GPURUN_EXTRA -- "$cmd" > /tmp/gpurun_$tag.log 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes" /tmp/gpurun_$tag.log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo $rc > /tmp/gpurun_$tag.done
) &
