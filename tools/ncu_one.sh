#!/bin/bash
# usage (on the GPU box): bash tools/ncu_one.sh <tag> <kernel regex> <mangled substring> [skip] [extra env...]
# one `ncu --set full` capture of one kernel of the bench step, summarised on the box (metrics + source-line stalls)
tag=$1; k=$2; m=$3; skip=${4:-3}
out=gpurun_out
mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o $out/${tag}_$k \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --direct-steps 0 > $out/${tag}_ncu_$k.log 2>&1
python tools/ncu_summary.py $out/${tag}_$k.ncu-rep > $out/${tag}_ncu_$k.md 2>&1
python tools/ncu_lines.py $out/${tag}_$k.ncu-rep monortm_b200/lib/libmonortm_b200.so $m 45 > $out/${tag}_lines_$k.txt 2>&1
rm -f $out/${tag}_$k.ncu-rep
