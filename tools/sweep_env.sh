# usage: bash tools/sweep_env.sh name1 "ENV=.. ENV=.." name2 "..."  -> short bench line per variant
while [ $# -gt 1 ]; do name=$1; envs=$2; shift 2
  env $envs timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --direct-steps 0 2> gpurun_out/sw_$name.err | tail -1 > gpurun_out/sw_$name.json
  echo "== $name: $(python -c "import json;d=json.load(open('gpurun_out/sw_$name.json'));print('ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['ms_per_step'],4),'lines',round(d['roofline']['kernel_ms'],4))" 2>&1)"
done
