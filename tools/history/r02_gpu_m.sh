#!/bin/bash
# Round 2, GPU call M (1 GPU): which commit cost the GPT-2 transposing load 10 % (0.146 ms at call E, 0.161 ms at HEAD, both static scheduling)?
# Libraries built from the commits between them, same box, same script.
#   /usr/local/graft/bin/gpurun --timeout 400 -- 'bash tools/r02/gpu_m.sh'
O=gpurun_out/r02m; mkdir -p $O
for v in c_ecf8bc4 c_7428869 c_d558897 default c_ecf8bc4 default; do
  lib=kukeon_b200/variants/libkukeon_gpuload.$v.so; [ $v = default ] && lib=kukeon_b200/libkukeon_gpuload.so
  [ -f $lib ] || { echo "$lib missing"; continue; }
  KUKEON_GPULOAD_LIB=$PWD/$lib KUKEON_GPULOAD_SCHED=static KK_QUICK_OUT=r02m/gpt2_$v.json timeout 90 python tools/gpu_quick_gpt2.py 2>$O/err_$v.txt | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'transposed', round(d['transposed_bf16']['ms_median'], 4), 'min', round(d['transposed_bf16']['ms_min'], 4), 'keep_f32', round(d['transposed_keep_f32']['ms_median'], 4), 'plain', round(d['plain_cast_bf16']['ms_median'], 4))
except Exception as e:
    print('$v', 'failed', e)"
done
echo "== done"
