#!/bin/bash
# One GPU call: C4 ensemble (tools/bench_configs.py --nprof 64 --configs c4) under a list of environment settings, each with the
# per-kernel launch summary of one call.   bash tools/sweep_env4.sh <tag> "ENV=a" "ENV=b ENV2=c" ...
tag=$1; shift 1
out=gpurun_out; mkdir -p $out
i=0
for v in "$@"; do
  i=$((i+1))
  echo "== $v $(env $v timeout 300 python tools/bench_configs.py --nprof 64 --configs c4 --reps 2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['config'], round(d['s_per_call']*1e3,2),'ms lines',round(d['last_call_kernel_ms']['lines'],2), end=' | ')
    except Exception: print(l.strip()[:200])
")"
  env $v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_$i.csv \
      python tools/bench_configs.py --nprof 64 --configs c4 --reps 1 > /dev/null 2>&1
  python tools/launch_summary.py $out/${tag}_$i.csv | grep mrtm | awk -F'|' '$4+0 > 300 {printf "   %s n %s total us %s\n",$2,$3,$4}'
done 2>&1 | tee $out/${tag}.log
