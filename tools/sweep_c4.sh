#!/bin/bash
# Build kernel variants (-D flags) on the GPU box and time each on the C4 ensemble (64 profiles) + C5.
#   bash tools/sweep_c4.sh name1 "-DA=1" name2 "-DB=2" ...
out=gpurun_out; mkdir -p $out
cd monortm_b200/csrc
names=()
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  names+=($name)
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -shared $flags \
      -o /tmp/lib_$name.so mrtm_api.cu mrtm_stage.cpp mrtm_host.cpp host/mrtm_driver.cpp > ../../$out/sw_${name}_build.log 2>&1 &
done
wait
cd ../..
for name in "${names[@]}"; do
  echo "== $name $(MRTM_LIB=/tmp/lib_$name.so $SWEEP_ENV timeout 300 python tools/bench_configs.py --nprof 64 --configs c4,c5,c2 --reps 2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['config'], round(d['s_per_call']*1e3,2),'ms lines',round(d['last_call_kernel_ms']['lines'],2), end=' | ')
    except Exception: print(l.strip()[:200])
")"
done
