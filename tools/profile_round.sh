#!/bin/bash
# One GPU call: bench at the default workload, the ncu launch list of the same command, and one
# `ncu --set full` capture per hot kernel.  Usage (from the repo root, on the GPU box):
#   bash tools/profile_round.sh <tag>        -> gpurun_out/<tag>_*
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
timeout 600 python bench.py --steps 20 --warmup 3 2> $out/${tag}_bench.err | tail -1 > $out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --direct-steps 0 > $out/${tag}_launches.log 2>&1
for k in near_kernel far_kernel derive_kernel voigt_kernel rt_kernel final_kernel plan_kernel; do
  skip=3
  [ $k = far_kernel ] && skip=11      # the level-0 launch of the 4th step (3 levels per step)
  [ $k = plan_kernel ] && skip=11
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o $out/${tag}_$k \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --direct-steps 0 > $out/${tag}_ncu_$k.log 2>&1
done
ls -la $out | tail -20
