for r in 8 7 6; do for w in 8 6 5; do
  MRTM_FF_RATIO=$r MRTM_FFW_RATIO=$w TUNE_STEPS=6 timeout 200 python tools/tune_ff.py 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().split('\n')[-1]);print('R',$r,'W',$w,'ms',round(d['ms_per_step'],4),'lines',round(d['lines_ms'],4),'acc',d['max_rel_far_vs_direct'],'far',d['far_expansions'],'direct',d['direct_evals'])"
done; done
