#!/bin/bash
# Build kernel variants (-D flags) on the GPU box and time each with a short bench + launch list.
#   bash tools/sweep_build.sh name1 "-DA=1 -DB=2" name2 "-DC=3" ...
out=gpurun_out; mkdir -p $out
cd monortm_b200/csrc
pids=()
names=()
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  names+=($name)
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -shared $flags \
      -o /tmp/lib_$name.so mrtm_api.cu mrtm_stage.cpp mrtm_host.cpp host/mrtm_driver.cpp > ../../$out/sw_${name}_build.log 2>&1 &
  pids+=($!)
done
wait
cd ../..
for name in "${names[@]}"; do
  case $name in *_F2) export MRTM_LINES_F=2;; *) unset MRTM_LINES_F;; esac
  MRTM_LIB=/tmp/lib_$name.so timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --direct-steps 0 2> $out/sw_$name.err | tail -1 > $out/sw_$name.json
  MRTM_LIB=/tmp/lib_$name.so timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 70 --csv --log-file $out/sw_${name}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --direct-steps 0 > /dev/null 2>&1
  echo "== $name: $(python -c "import json;d=json.load(open('$out/sw_$name.json'));print('ms/step',round(d['ms_per_step'],4),'lines',round(d['roofline']['kernel_ms'],4))" 2>&1)"
  python tools/launch_summary.py $out/sw_${name}_launches.csv | grep -E "near2_kernel<., 0|near_kernel<., 0|far_kernel|voigtT?_kernel<|final_kernel<|derive|rt_kernel|plan_kernel" | awk -F'|' '{printf "   %s avg %s\n",$2,$5}'
done
