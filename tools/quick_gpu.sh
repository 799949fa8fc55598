#!/bin/bash
# One GPU call while iterating on a kernel: parity tests, a short bench line and the per-kernel launch list.
#   bash tools/quick_gpu.sh <tag> [notest]
tag=${1:-q}
out=gpurun_out
mkdir -p $out
if [ "$2" != "notest" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/${tag}_pytest.log; fi
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --direct-steps 0 2> $out/${tag}_bench.err | tail -1 > $out/${tag}_bench.json
python -c "import json;d=json.load(open('$out/${tag}_bench.json'));print('ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'lines',d['roofline']['kernel_ms'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --direct-steps 0 > $out/${tag}_launches.log 2>&1
python tools/launch_summary.py $out/${tag}_launches.csv | tee $out/${tag}_launches.md
