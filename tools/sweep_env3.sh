#!/bin/bash
# One GPU call: the coarse-list configs (C4 ensemble, C5, C2; tools/bench_configs.py --nprof 64) under a list of environment settings,
# optionally with the GPU parity tests under each setting.   bash tools/sweep_env3.sh <tag> "<tests or empty>" "ENV=a" "ENV=b ENV2=c" ...
tag=$1; tests=$2; shift 2
out=gpurun_out; mkdir -p $out
for v in "$@"; do
  if [ -n "$tests" ]; then echo "-- tests under $v: $(env $v timeout 900 python -m pytest $tests -m gpu -x -q 2>&1 | tail -3 | tr '\n' ' ')"; fi
  echo "== $v $(env $v timeout 300 python tools/bench_configs.py --nprof 64 --configs c4,c5,c2 --reps 2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['config'], round(d['s_per_call']*1e3,2),'ms lines',round(d['last_call_kernel_ms']['lines'],2), end=' | ')
    except Exception: print(l.strip()[:200])
")"
done 2>&1 | tee $out/${tag}.log
