#!/bin/bash
# Round 2, GPU call H (1 GPU): Q4_K tile-size sweep.  Q5_K runs the same quad code at 0.96-0.98 of the copy peak (186-block tiles), Q4_K at 0.87
# (224-block tiles) with FEWER instructions and LESS traffic per weight: something about the tile geometry, not the arithmetic.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/r02/gpu_h.sh'
O=gpurun_out/r02h; mkdir -p $O
for v in default q186 q200 q208 q216 q221 q223 default; do
  lib=kukeon_b200/variants/libkukeon_gpuload.$v.so; [ $v = default ] && lib=kukeon_b200/libkukeon_gpuload.so
  [ -f $lib ] || { echo "$lib missing"; continue; }
  KUKEON_GPULOAD_LIB=$PWD/$lib KK_QUICK_OUT=r02h/q4k_$v.json timeout 120 python tools/gpu_quick_types.py --types Q4_K --weights-m 4096 --passes 20 > $O/q4k_$v.stdout 2>&1; echo "$v rc=$?"
  grep -E "^Q4_K " $O/q4k_$v.stdout | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l.split(' ', 1)[1]); print('   ms_median', round(d['ms_median'], 4), 'min', round(d['ms_min'], 4), 'frac', round(d['frac_of_copy_peak'], 4))
"
done
echo "== done"
