#!/bin/bash
# First GPU call of the next round (1 GPU, ~5 minutes of box time): everything written after round 1's budget ran out, in the order
# "cheapest and most informative first", each step under its own timeout so a hang cannot eat the call.  Results land in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 780 -- 'bash tools/gpu_next_first.sh'
mkdir -p gpurun_out
echo "== 1. torch-free parity of the new paths (rewritten dequantisers, wide-store tiles, NVLS is multi-GPU and not here)"
timeout 60 python tools/gpu_quick.py > gpurun_out/next_quick.stdout 2>&1; echo "rc=$?"; tail -25 gpurun_out/quick.log
echo "== 2. transpose geometry A/B/C on GPT-2-small"
timeout 60 python tools/gpu_quick_t8.py > gpurun_out/next_t8.stdout 2>&1; echo "rc=$?"; tail -c 1200 gpurun_out/t8_ab.json
echo "== 3. roofline table of every dequantiser"
timeout 150 python tools/gpu_quick_types.py --weights-m 512 > gpurun_out/next_types.stdout 2>&1; echo "rc=$?"; tail -30 gpurun_out/next_types.stdout | cut -c1-220
echo "== 4. the queued parity file (skips the multi-GPU cases on one GPU)"
timeout 200 python -m pytest tests/test_zz_gpu_quants_f4.py -x -q -m gpu -p no:cacheprovider > gpurun_out/next_zz_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/next_zz_pytest.log | cut -c1-220
echo "== 5. A/B builds (make -C kukeon_b200/csrc variants; loaded with KUKEON_GPULOAD_LIB): producer on kk_make_tile (ps1), warp-cooperative row copies (ps2), equal-width 8-row tiles (t8bal), 20 consumer warps at 80 registers (cw20)"
for v in ps1 ps2 t8bal cw20; do
  lib=kukeon_b200/variants/libkukeon_gpuload.$v.so
  [ -f $lib ] || { echo "$lib missing (make -C kukeon_b200/csrc variants)"; continue; }
  KUKEON_GPULOAD_LIB=$PWD/$lib timeout 60 python tools/gpu_quick.py > gpurun_out/next_quick_$v.stdout 2>&1; echo "$v parity rc=$?"
  KUKEON_GPULOAD_LIB=$PWD/$lib KK_QUICK_OUT=t8_ab_$v.json timeout 60 python tools/gpu_quick_t8.py > gpurun_out/next_t8_$v.stdout 2>&1; echo "$v t8 rc=$?"; tail -c 600 gpurun_out/t8_ab_$v.json
done
echo "== 6. A/B builds of the Q4_K tile geometry (q4k192: 192-block tiles; q4kbal: 14 contiguous blocks per warp) against the default"
for v in default q4k192 q4kbal q4krot cw20; do
  lib=kukeon_b200/variants/libkukeon_gpuload.$v.so; [ $v = default ] && lib=kukeon_b200/libkukeon_gpuload.so
  [ -f $lib ] || { echo "$lib missing"; continue; }
  KUKEON_GPULOAD_LIB=$PWD/$lib KK_QUICK_OUT=q4k_$v.json timeout 90 python tools/gpu_quick_types.py --types Q4_K --weights-m 2048 --passes 20 > gpurun_out/next_q4k_$v.stdout 2>&1; echo "$v rc=$?"; tail -3 gpurun_out/next_q4k_$v.stdout | cut -c1-200
done
