#!/bin/bash
# Round 2, GPU call E (1 GPU): the HEAD kernel (FMA form + warp vote for Q4_K, Q5_K quads, conflict-free F32 cast, VMM pools) — full suite,
# per-type table, HEAD ncu captures of the ops that changed, sanitizer over them, the bench line of both arms, full-size Mixtral q4_K.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r02/gpu_e.sh'
O=gpurun_out/r02e; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on -k regex:kk_convert_kernel"
echo "== 1. torch-free parity"
timeout 90 python tools/gpu_quick.py > $O/quick.stdout 2>&1; echo "rc=$?"; grep -c PASS $O/quick.stdout; grep -v PASS $O/quick.stdout | tail -4 | cut -c1-200
echo "== 2. full GPU suite (verbose: every test name in the log)"
timeout 700 python -m pytest tests/ -v -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.log | cut -c1-240; grep -c PASSED $O/pytest_gpu.log; grep -E "FAILED|ERROR" $O/pytest_gpu.log | head
echo "== 3. roofline table of every dequantiser (1 G weights each)"
KK_QUICK_OUT=r02e/types_roofline.json timeout 400 python tools/gpu_quick_types.py --weights-m 1024 --passes 10 > $O/types.stdout 2>&1; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02e/types_roofline.json"))
    for k, v in d["types"].items():
        print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a in ("ms_median", "GBps", "frac_of_copy_peak", "write_GBps", "error")})
except Exception as e:
    print("no table:", e)
PY
echo "== 4. GPT-2 transposing load; Q4_K / Q5_K at 4 G weights"
KK_QUICK_OUT=r02e/gpt2_quick.json timeout 120 python tools/gpu_quick_gpt2.py > $O/gpt2.stdout 2>&1; echo "rc=$?"; tail -c 900 $O/gpt2.stdout; echo
KK_QUICK_OUT=r02e/q45k_4g.json timeout 200 python tools/gpu_quick_types.py --types Q4_K,Q5_K --weights-m 4096 --passes 20 > $O/q45k_4g.stdout 2>&1; echo "rc=$?"; grep -E "^Q[45]_K" $O/q45k_4g.stdout | cut -c1-260
echo "== 5. ncu --set full at HEAD: Q4_K, Q5_K, F32 cast, GPT-2"
for T in Q4_K Q5_K F32; do
  timeout 240 $NCU -s 3 -c 1 -o $O/prof_$T -f python tools/gpu_quick_types.py --types $T --weights-m 1024 --passes 1 > $O/ncu_$T.log 2>&1; echo "ncu $T rc=$?"
done
timeout 240 $NCU -s 3 -c 1 -o $O/prof_gpt2 -f python tools/gpu_quick_gpt2.py > $O/ncu_gpt2.log 2>&1; echo "ncu gpt2 rc=$?"
for f in $O/prof_*.ncu-rep; do b=${f%.ncu-rep}; ncu -i $f --page raw --csv > $b.raw.csv 2>/dev/null; ncu -i $f --page details > $b.details.txt 2>/dev/null; done
rm -f $O/prof_F32.ncu-rep
echo "== 6. compute-sanitizer over the changed ops"
SAN=/usr/local/cuda/bin/compute-sanitizer
K="golden_q4k_values or q4_k_m_style or mixed_safetensors_every_op or gpt2_conv1d_transpose or mixtral_style_gguf_q4k"
timeout 300 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_load.py -q -m gpu -p no:cacheprovider -k "$K" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitizer_memcheck.log | tail -3
timeout 300 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_load.py -q -m gpu -p no:cacheprovider -k "$K" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/sanitizer_racecheck.log | tail -3
timeout 200 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_quants.py -q -m gpu -p no:cacheprovider -k "golden_fixture_values or llama_shaped_mix" > $O/sanitizer_memcheck_quants.log 2>&1; echo "memcheck quants rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitizer_memcheck_quants.log | tail -3
echo "== 7. the bench line (both arms)"
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "ours rc=$?"; head -c 600 $O/bench_n1.json; echo
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; head -c 400 $O/bench_ref.json; echo
echo "== 8. full-size Mixtral-8x7B q4_K on one GPU (26.3 GB in, 93.4 GB out)"
timeout 500 python bench.py --workload mixtral-q4k --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_mixtral_full_n1.json 2> $O/bench_mixtral_full_n1.err; echo "rc=$?"; head -c 1500 $O/bench_mixtral_full_n1.json; echo; tail -2 $O/bench_mixtral_full_n1.err
echo "== done"
