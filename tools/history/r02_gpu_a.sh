#!/bin/bash
# Round 2, GPU call A (1 GPU): green suite first, then the measurements that pick winners among the round-1 A/B builds.
#   /usr/local/graft/bin/gpurun --timeout 1000 -- 'bash tools/r02/gpu_a.sh'
# Every step has its own timeout; results land in gpurun_out/r02a/.
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/box.txt 2>&1; nproc >> $O/box.txt
echo "== 1. torch-free parity of every path (default build)"
timeout 90 python tools/gpu_quick.py > $O/quick.stdout 2>&1; echo "rc=$?"; tail -5 $O/quick.stdout | cut -c1-200
echo "== 2. full GPU suite, no -x (every failure is reported, none masks another)"
timeout 500 python -m pytest tests/ -q -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 $O/pytest_gpu.log | cut -c1-240
echo "== 3. roofline table of every dequantiser (resident image, event-timed in the library)"
timeout 200 python tools/gpu_quick_types.py --weights-m 512 --passes 8 > $O/types.stdout 2>&1; echo "rc=$?"; cp gpurun_out/types_roofline.json $O/ 2>/dev/null
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02a/types_roofline.json"))
    for k, v in d["types"].items():
        print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a in ("ms_median", "GBps", "frac_of_copy_peak", "write_GBps", "error")})
except Exception as e:
    print("no table:", e)
PY
echo "== 4. transpose geometry A/B/C on GPT-2-small: default build, then ps1 / ps2 / t8bal"
timeout 90 python tools/gpu_quick_t8.py > $O/t8_default.stdout 2>&1; echo "default rc=$?"; cp gpurun_out/t8_ab.json $O/t8_ab_default.json; tail -c 900 $O/t8_ab_default.json
for v in ps1 ps2 t8bal cw20; do
  lib=kukeon_b200/variants/libkukeon_gpuload.$v.so
  [ -f $lib ] || { echo "$lib missing"; continue; }
  KUKEON_GPULOAD_LIB=$PWD/$lib timeout 90 python tools/gpu_quick.py > $O/quick_$v.stdout 2>&1; echo "$v parity rc=$?"; tail -2 $O/quick_$v.stdout | cut -c1-200
  KUKEON_GPULOAD_LIB=$PWD/$lib KK_QUICK_OUT=r02a/t8_ab_$v.json timeout 90 python tools/gpu_quick_t8.py > $O/t8_$v.stdout 2>&1; echo "$v t8 rc=$?"; tail -c 700 $O/t8_ab_$v.json
done
echo "== 5. Q4_K tiling A/B (4 G weights, 20 passes) + copy/cast for the producer variant"
for v in default q4k192 q4kbal q4krot cw20 ps1; do
  lib=kukeon_b200/variants/libkukeon_gpuload.$v.so; [ $v = default ] && lib=kukeon_b200/libkukeon_gpuload.so
  [ -f $lib ] || { echo "$lib missing"; continue; }
  KUKEON_GPULOAD_LIB=$PWD/$lib KK_QUICK_OUT=r02a/q4k_$v.json timeout 120 python tools/gpu_quick_types.py --types Q4_K --weights-m 4096 --passes 20 > $O/q4k_$v.stdout 2>&1; echo "$v rc=$?"
  grep -E "^Q4_K " $O/q4k_$v.stdout | cut -c1-230
done
for v in default ps1 cw20; do
  lib=kukeon_b200/variants/libkukeon_gpuload.$v.so; [ $v = default ] && lib=kukeon_b200/libkukeon_gpuload.so
  KUKEON_GPULOAD_LIB=$PWD/$lib KK_QUICK_OUT=r02a/casts_$v.json timeout 120 python tools/gpu_quick_types.py --types BF16,F16,F32,Q6_K,Q8_0 --weights-m 1024 --passes 10 > $O/casts_$v.stdout 2>&1; echo "$v rc=$?"
  grep -E "^(BF16|F16|F32|Q6_K|Q8_0) " $O/casts_$v.stdout | cut -c1-230
done
echo "== 6. compute-sanitizer memcheck / racecheck over the core parity cases"
SAN=/usr/local/cuda/bin/compute-sanitizer
K="mixed_safetensors_every_op or golden_files or gpt2_conv1d_transpose or q4_k_m_style"
timeout 240 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_load.py -q -m gpu -p no:cacheprovider -k "$K" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitizer_memcheck.log | tail -3
timeout 300 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_load.py -q -m gpu -p no:cacheprovider -k "$K" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/sanitizer_racecheck.log | tail -3
echo "== done"
