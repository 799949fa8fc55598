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
for k in near2_kernel far_warp_kernel far_kernel derive_kernel voigtT_kernel rt_kernel final_kernel plan_kernel; do
  skip=3
  [ $k = far_warp_kernel ] && skip=7  # the level-0 launch of the 4th step (two far_warp levels per step)
  [ $k = plan_kernel ] && skip=11
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o $out/${tag}_$k \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --direct-steps 0 > $out/${tag}_ncu_$k.log 2>&1
done
ls -la $out | tail -20
# reports are ~9 MB each and gpurun brings back at most 64 MiB: summarise on the box, keep only the summaries
python tools/ncu_summary.py $out/${tag}_*.ncu-rep > $out/${tag}_ncu_summary.md 2>&1
for k in near2_kernel far_warp_kernel voigtT_kernel final_kernel derive_kernel rt_kernel; do
  m=$k      # mangled-name substring of the instantiation that was captured (templates have several)
  [ $k = near2_kernel ] && m=near2_kernelILi4ELb0E
  [ $k = voigtT_kernel ] && m=voigtT_kernelILi4E
  [ $k = final_kernel ] && m=final_kernelILi4E
  python tools/ncu_lines.py $out/${tag}_$k.ncu-rep monortm_b200/lib/libmonortm_b200.so $m 40 > $out/${tag}_lines_$k.txt 2>&1
done
rm -f $out/${tag}_*.ncu-rep
