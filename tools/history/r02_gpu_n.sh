#!/bin/bash
# Round 2, GPU call N (1 GPU): the two-instantiation kernel (static loops for transposing plans, dynamic draws for everything else): GPT-2 timing,
# full suite, sanitizer on both instantiations, both bench arms.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r02/gpu_n.sh'
O=gpurun_out/r02n; mkdir -p $O
timeout 90 python tools/gpu_quick.py > $O/quick.stdout 2>&1; echo "parity rc=$? pass=$(grep -c PASS $O/quick.stdout) fail=$(grep -c FAIL $O/quick.stdout)"
KK_QUICK_OUT=r02n/gpt2_quick.json timeout 120 python tools/gpu_quick_gpt2.py > $O/gpt2.stdout 2>&1; echo "rc=$?"; tail -c 700 $O/gpt2.stdout; echo
KK_QUICK_OUT=r02n/types.json timeout 200 python tools/gpu_quick_types.py --types BF16,Q4_K,Q8_0,F32 --weights-m 1024 --passes 10 2>/dev/null | grep -E "^(BF16|Q4_K|Q8_0|F32) " | cut -c1-120
timeout 700 python -m pytest tests/ -q -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log | cut -c1-200
SAN=/usr/local/cuda/bin/compute-sanitizer
K="mixed_safetensors_every_op or gpt2_conv1d_transpose or q4_k_m_style or golden_files"
timeout 300 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_load.py -q -m gpu -p no:cacheprovider -k "$K" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/sanitizer_racecheck.log | tail -2
timeout 300 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_load.py -q -m gpu -p no:cacheprovider -k "$K" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitizer_memcheck.log | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "ours rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02n/bench_n1.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "time_to_agent_ready_s")}, d["roofline"]["frac"])
    print("secondary", d["secondary"]["roofline"]["frac"], d["secondary"]["roofline"]["hbm_write_frac"], "gpt2", d["secondary_gpt2"]["ms_per_step"], d["secondary_gpt2"]["roofline"]["frac"])
    print("e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
    r = json.loads(open("gpurun_out/r02n/bench_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["ms_per_step"])
except Exception as e:
    print("unreadable", e)
PY
echo "== done"
