#!/bin/bash
# One GPU call: A/B of an environment switch on the default bench workload (+ optional parity tests and launch list).
#   bash tools/quick_ab.sh <tag> "<ENV=a>" "<ENV=b>" [tests]
tag=$1; A=$2; B=$3
out=gpurun_out
mkdir -p $out
if [ -n "$4" ]; then timeout 900 python -m pytest $4 -m gpu -x -q 2>&1 | tail -8 | tee $out/${tag}_pytest.log; fi
for v in "$A" "$B"; do
  n=$(echo "$v" | tr -c 'A-Za-z0-9=\n' '_')
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --direct-steps 0 2> $out/${tag}_${n}.err | tail -1 > $out/${tag}_${n}.json
  python -c "import json;d=json.load(open('$out/${tag}_${n}.json'));print('$v','ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'lines',d['roofline']['kernel_ms'],'parity',d.get('parity_check',{}).get('max_rel_od'),d.get('parity_check',{}).get('pass'))" | tee -a $out/${tag}_ab.log
done
env $B timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --direct-steps 0 > $out/${tag}_launches.log 2>&1
python tools/launch_summary.py $out/${tag}_launches.csv | tee $out/${tag}_launches.md
