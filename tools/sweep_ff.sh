run() { # name env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --nwn-per-gpu 125000 --steps 8 --warmup 3 --no-cpu-baseline --direct-steps 0 $EXTRA > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err
}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s3_pytest13.log
run base X=1
run L2 MRTM_FF_LEVELS=2
run L3S4 MRTM_FF_S=4
run L3S16 MRTM_FF_S=16
run L2S16 MRTM_FF_LEVELS=2 MRTM_FF_S=16
EXTRA="--n-filler 4096"
run fast_base X=1
EXTRA=""
ncu --metrics gpu__time_duration.sum --clock-control none -s 44 -c 14 --csv --log-file gpurun_out/launches_v6f.csv python bench.py --nwn-per-gpu 125000 --steps 2 --warmup 3 --no-cpu-baseline --direct-steps 0 > gpurun_out/ncu_v6b.log 2>&1
