#!/bin/bash
# One GPU call: the C4 ensemble (bench.py --config c4, 64 profiles) -- launch list and one `ncu --set full` capture of near_kernel and far_warp_kernel.
#   bash tools/profile_c4.sh <tag>
tag=${1:-c4}
out=gpurun_out; mkdir -p $out
cmd="python bench.py --config c4 --nprof-per-gpu 64 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches.csv $cmd > $out/${tag}_launches.log 2>&1
python tools/launch_summary.py $out/${tag}_launches.csv > $out/${tag}_launches.md
for k in near_kernel far_warp_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $out/${tag}_$k $cmd > $out/${tag}_ncu_$k.log 2>&1
done
python tools/ncu_summary.py $out/${tag}_*.ncu-rep > $out/${tag}_ncu_summary.md 2>&1
python tools/ncu_lines.py $out/${tag}_near_kernel.ncu-rep monortm_b200/lib/libmonortm_b200.so near_kernelILi1ELb0ELi32 30 > $out/${tag}_lines_near_kernel.txt 2>&1
rm -f $out/${tag}_*.ncu-rep
