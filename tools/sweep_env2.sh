#!/bin/bash
# time the bench (with the oracle parity check) under environment variants:  bash tools/sweep_env2.sh "A=1" "B=2 C=3" ...
for v in "$@"; do
  echo "== $v: $(env $v timeout 300 python bench.py --steps 10 --warmup 3 --direct-steps 0 --cpu-sample-nwn 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['parity_check']; print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), 'od', p['max_rel_od'], 'tb', p['max_dtb_K'], p['sel_exact'], p['pass'])")"
done
