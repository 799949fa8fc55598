#!/bin/bash
# far-field parameter sweep (R = MRTM_FF_RATIO, K = Taylor terms via library variant, F = frequencies per thread, S = level fan-out)
L=monortm_b200/lib
out=gpurun_out/sweep_ff2.jsonl
: > $out
run() { env "$@" timeout 200 python tools/tune_ff.py >> $out 2>> gpurun_out/sweep_ff2.err; }
run X=1
run MRTM_FF_RATIO=6
run MRTM_FF_RATIO=5
run MRTM_LINES_F=2
run MRTM_LINES_F=2 MRTM_FF_RATIO=6
run MRTM_FF_RATIO=6 MRTM_FF_S=4
run MRTM_FF_RATIO=6 MRTM_FF_S=16
run MRTM_LIB=$PWD/$L/libmonortm_b200_k16.so MRTM_FF_RATIO=6
run MRTM_LIB=$PWD/$L/libmonortm_b200_k16.so MRTM_FF_RATIO=5
run MRTM_LIB=$PWD/$L/libmonortm_b200_k16.so MRTM_FF_RATIO=4
run MRTM_LIB=$PWD/$L/libmonortm_b200_k16.so MRTM_FF_RATIO=5 MRTM_LINES_F=2
run MRTM_LIB=$PWD/$L/libmonortm_b200_k16.so MRTM_FF_RATIO=4 MRTM_LINES_F=2
run MRTM_LIB=$PWD/$L/libmonortm_b200_k12.so MRTM_FF_RATIO=8
run MRTM_LIB=$PWD/$L/libmonortm_b200_k12.so MRTM_FF_RATIO=10
cat $out | cut -c1-400
tail -5 gpurun_out/sweep_ff2.err
