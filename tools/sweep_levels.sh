run() { name=$1; shift
  env "$@" timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --direct-steps 0 2> gpurun_out/sw_$name.err | tail -1 > gpurun_out/sw_$name.json
  echo "== $name: $(python -c "import json;d=json.load(open('gpurun_out/sw_$name.json'));print('ms/step',round(d['ms_per_step'],4),'lines',round(d['roofline']['kernel_ms'],4),'far',d['roofline']['far_expansions_per_launch'],'direct',d['roofline']['direct_evals_per_launch'])" 2>&1)"
}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run base X=1
run nowarp MRTM_FARW_MIN=100000000
run L3S8w1000 MRTM_FARW_MIN=1000
run L4S4 MRTM_FF_LEVELS=4 MRTM_FF_S=4
run L4S4w1000 MRTM_FF_LEVELS=4 MRTM_FF_S=4 MRTM_FARW_MIN=1000
run L3S6 MRTM_FF_LEVELS=3 MRTM_FF_S=6
run L3S6w1000 MRTM_FF_LEVELS=3 MRTM_FF_S=6 MRTM_FARW_MIN=1000
run L4S5 MRTM_FF_LEVELS=4 MRTM_FF_S=5
