#!/bin/bash
# Round 2, GPU call H (1 GPU): dynamic tile scheduling (CTAs draw tile batches from a global counter after their first, static tile).  ncu had shown
# the SMs active for only 0.86 (Q4_K) / 0.92 (bf16 copy) of the kernel's duration under static round-robin: the kernel ended with its slowest CTA.
# Then the Q4_K tile-size sweep on the new scheduler.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/r02/gpu_h.sh'
O=gpurun_out/r02h; mkdir -p $O
echo "== 1. parity"
timeout 90 python tools/gpu_quick.py > $O/quick.stdout 2>&1; echo "rc=$?"; grep -c PASS $O/quick.stdout; grep -v PASS $O/quick.stdout | tail -3 | cut -c1-200
echo "== 2. per-type table"
KK_QUICK_OUT=r02h/types_roofline.json timeout 400 python tools/gpu_quick_types.py --weights-m 1024 --passes 10 > $O/types.stdout 2>&1; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02h/types_roofline.json"))
    print({k: round(v.get("frac_of_copy_peak", 0), 3) for k, v in d["types"].items()})
except Exception as e:
    print("no table:", e)
PY
echo "== 3. 4 Gi weights: Q4_K, Q5_K, BF16; GPT-2"
KK_QUICK_OUT=r02h/q45k_4g.json timeout 200 python tools/gpu_quick_types.py --types Q4_K,Q5_K,BF16 --weights-m 4096 --passes 20 > $O/q45k_4g.stdout 2>&1; echo "rc=$?"; grep -E "^(Q[45]_K|BF16) " $O/q45k_4g.stdout | cut -c1-200
KK_QUICK_OUT=r02h/gpt2_quick.json timeout 120 python tools/gpu_quick_gpt2.py > $O/gpt2.stdout 2>&1; echo "rc=$?"; tail -c 700 $O/gpt2.stdout; echo
echo "== 4. Q4_K tile-size sweep"
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
echo "== 5. full GPU suite"
timeout 700 python -m pytest tests/ -q -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.log | cut -c1-240
echo "== 6. racecheck / memcheck on the new scheduler"
SAN=/usr/local/cuda/bin/compute-sanitizer
K="mixed_safetensors_every_op or golden_files or gpt2_conv1d_transpose or q4_k_m_style or multi_destination_store_paths_on_one_gpu"
timeout 300 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_load.py -q -m gpu -p no:cacheprovider -k "$K" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/sanitizer_racecheck.log | tail -3
timeout 300 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_load.py -q -m gpu -p no:cacheprovider -k "$K" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitizer_memcheck.log | tail -3
echo "== 7. bench line, ncu of the copy launch"
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "ours rc=$?"; head -c 400 $O/bench_n1.json; echo; tail -3 $O/bench_n1.err
timeout 240 ncu --set full --clock-control none --import-source on -k regex:kk_convert_kernel -s 3 -c 1 -o $O/prof_BF16 -f python tools/gpu_quick_types.py --types BF16 --weights-m 1024 --passes 1 > $O/ncu_BF16.log 2>&1; echo "ncu rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:kk_convert_kernel -s 3 -c 1 -o $O/prof_Q4_K -f python tools/gpu_quick_types.py --types Q4_K --weights-m 1024 --passes 1 > $O/ncu_Q4_K.log 2>&1; echo "ncu rc=$?"
for f in $O/prof_*.ncu-rep; do b=${f%.ncu-rep}; ncu -i $f --page raw --csv > $b.raw.csv 2>/dev/null; ncu -i $f --page details > $b.details.txt 2>/dev/null; rm -f $f; done
echo "== done"
