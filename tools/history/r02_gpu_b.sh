#!/bin/bash
# Round 2, GPU call B (1 GPU): first hardware run of the round-2 kernel (plain shared loads + loads hoisted / unrolled, one transpose family, producer on
# kk_make_tile), the per-type roofline table, ncu --set full of the copy / Q4_K / slowest dequantisers / GPT-2 launches, sanitizer on the remaining
# verdict-listed cases, the new bench line, and the host-register probe for the e2e leg.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r02/gpu_b.sh'
O=gpurun_out/r02b; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on -k regex:kk_convert_kernel"
echo "== 1. torch-free parity of every path"
timeout 90 python tools/gpu_quick.py > $O/quick.stdout 2>&1; echo "rc=$?"; grep -c PASS $O/quick.stdout; grep -v PASS $O/quick.stdout | tail -5 | cut -c1-200
echo "== 2. roofline table of every dequantiser (1 G weights each)"
KK_QUICK_OUT=r02b/types_roofline.json timeout 400 python tools/gpu_quick_types.py --weights-m 1024 --passes 10 > $O/types.stdout 2>&1; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02b/types_roofline.json"))
    for k, v in d["types"].items():
        print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a in ("ms_median", "GBps", "frac_of_copy_peak", "write_GBps", "error")})
except Exception as e:
    print("no table:", e)
PY
echo "== 3. GPT-2 transposing load"
KK_QUICK_OUT=r02b/gpt2_quick.json timeout 120 python tools/gpu_quick_gpt2.py > $O/gpt2.stdout 2>&1; echo "rc=$?"; tail -c 1200 $O/gpt2.stdout
echo "== 4. Q4_K at 4 G weights"
KK_QUICK_OUT=r02b/q4k_4g.json timeout 120 python tools/gpu_quick_types.py --types Q4_K --weights-m 4096 --passes 20 > $O/q4k_4g.stdout 2>&1; echo "rc=$?"; grep "^Q4_K" $O/q4k_4g.stdout | cut -c1-250
echo "== 5. ncu --set full: one launch each (skip the 3 warm-ups)"
for T in BF16 Q4_K Q3_K Q5_K IQ2_XXS TQ1_0; do
  timeout 240 $NCU -s 3 -c 1 -o $O/prof_$T -f python tools/gpu_quick_types.py --types $T --weights-m 1024 --passes 1 > $O/ncu_$T.log 2>&1; echo "ncu $T rc=$?"
done
timeout 240 $NCU -s 3 -c 1 -o $O/prof_gpt2 -f python tools/gpu_quick_gpt2.py > $O/ncu_gpt2.log 2>&1; echo "ncu gpt2 rc=$?"
for f in $O/prof_*.ncu-rep; do
  b=${f%.ncu-rep}
  ncu -i $f --page raw --csv > $b.raw.csv 2>/dev/null
  ncu -i $f --page details > $b.details.txt 2>/dev/null
done
ls -la $O/*.ncu-rep | awk '{print $5, $9}'
echo "== 6. full GPU suite"
timeout 600 python -m pytest tests/ -q -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 $O/pytest_gpu.log | cut -c1-240
echo "== 7. compute-sanitizer on the remaining verdict-listed cases"
SAN=/usr/local/cuda/bin/compute-sanitizer
K="multi_destination_store_paths_on_one_gpu or virtual_ranks_scatter_exchange_on_one_gpu or gpt2_conv1d_transpose or mixed_safetensors_every_op or q4_k_m_style"
timeout 400 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_load.py -q -m gpu -p no:cacheprovider -k "$K" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitizer_memcheck.log | tail -3
timeout 400 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_load.py -q -m gpu -p no:cacheprovider -k "$K" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/sanitizer_racecheck.log | tail -3
echo "== 8. the bench line (both arms)"
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "ours rc=$?"; head -c 3000 $O/bench_n1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cat $O/bench_ref.json
echo "== 9. ncu launch list of the kernel-only bench"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_copy_resident.csv python bench.py --steps 2 --warmup 3 --kernel-only --no-cpu-baseline > $O/bench_under_ncu.log 2>&1; echo "rc=$?"; grep -c kk_convert $O/launches_copy_resident.csv
echo "== 10. host-register probe (e2e leg)"
timeout 200 python tools/hostreg_probe.py 4 > $O/hostreg_probe.json 2> $O/hostreg_probe.err; echo "rc=$?"; cat $O/hostreg_probe.json; tail -3 $O/hostreg_probe.err
echo "== done"
