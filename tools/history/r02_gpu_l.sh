#!/bin/bash
# Round 2, GPU call L (1 GPU): transposing tile shape x scheduler.  8-row tiles store 16 bytes per column and rely on the neighbouring row group's tile
# to fill the 32-byte sector in L2 (static scheduling keeps them close in time; dynamic draws measured 8 % slower).  16- and 32-row tiles store whole
# sectors / whole 64-byte runs per thread and should not care.
#   /usr/local/graft/bin/gpurun --timeout 500 -- 'bash tools/r02/gpu_l.sh'
O=gpurun_out/r02l; mkdir -p $O
for v in default t16 t32; do
  lib=kukeon_b200/variants/libkukeon_gpuload.$v.so; [ $v = default ] && lib=kukeon_b200/libkukeon_gpuload.so
  [ -f $lib ] || { echo "$lib missing"; continue; }
  KUKEON_GPULOAD_LIB=$PWD/$lib timeout 90 python tools/gpu_quick.py > $O/quick_$v.stdout 2>&1; echo "$v parity rc=$? pass=$(grep -c PASS $O/quick_$v.stdout) fail=$(grep -c FAIL $O/quick_$v.stdout)"
  for s in static dynamic static dynamic; do
    KUKEON_GPULOAD_LIB=$PWD/$lib KUKEON_GPULOAD_SCHED=$s KK_QUICK_OUT=r02l/gpt2_${v}_$s.json timeout 90 python tools/gpu_quick_gpt2.py 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '$s', 'transposed', round(d['transposed_bf16']['ms_median'], 4), round(d['transposed_bf16']['frac_of_copy_peak'], 3), 'keep_f32', round(d['transposed_keep_f32']['ms_median'], 4), 'plain', round(d['plain_cast_bf16']['ms_median'], 4))"
  done
done
echo "== done"
