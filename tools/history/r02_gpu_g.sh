#!/bin/bash
# Round 2, GPU call G (1 GPU): e2e leg tuning at N = 1 (>= 0.93 of the pinned-H2D probe; 0.86 at the defaults 16 readers x 32 slots x 16 MiB).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r02/gpu_g.sh'
O=gpurun_out/r02g; mkdir -p $O
echo "== 0. the warp-rotation build (every warp 4 + 3 Q4_K quads over two tiles): parity, per-type table, Q4_K / Q5_K at 4 Gi weights, core GPU tests"
timeout 90 python tools/gpu_quick.py > $O/quick.stdout 2>&1; echo "rc=$?"; grep -c PASS $O/quick.stdout; grep -v PASS $O/quick.stdout | tail -3 | cut -c1-200
KK_QUICK_OUT=r02g/types_roofline.json timeout 400 python tools/gpu_quick_types.py --weights-m 1024 --passes 10 > $O/types.stdout 2>&1; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02g/types_roofline.json"))
    print({k: round(v.get("frac_of_copy_peak", 0), 3) for k, v in d["types"].items()})
except Exception as e:
    print("no table:", e)
PY
KK_QUICK_OUT=r02g/q45k_4g.json timeout 200 python tools/gpu_quick_types.py --types Q4_K,Q5_K,Q6_K,Q8_0 --weights-m 4096 --passes 20 > $O/q45k_4g.stdout 2>&1; echo "rc=$?"; grep -E "^Q[4568]_[K0]" $O/q45k_4g.stdout | cut -c1-200
timeout 500 python -m pytest tests/test_gpu_load.py tests/test_gpu_quants.py -q -m gpu -p no:cacheprovider -x > $O/pytest_gpu_core.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_core.log | cut -c1-200
echo "== 1. e2e sweep"
: > $O/e2e_sweep.jsonl
run() {  # readers slots slot_mb [extra flags]
  r=$1; s=$2; mb=$3; shift 3
  timeout 150 python bench.py --e2e-only --steps 4 --warmup 2 --readers $r --slots $s --slot-mb $mb --keep-data "$@" 2> $O/e2e_last.err | tail -1 >> $O/e2e_sweep.jsonl
  tail -1 $O/e2e_sweep.jsonl | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('readers', d['config']['readers'], 'slots', d['config']['slots'], 'slot_mb', d['config']['slot_mb'], 'zerocopy', d['config']['zerocopy'], '->', round(d['e2e']['value'], 2), 'GB/s', round(d['e2e']['ms_per_step'], 1), 'ms; probe', round(d['h2d_probe_GBps'] or 0, 1))
except Exception as e:
    print('unreadable', e)
"
}
run 16 32 16
run 16 48 16
run 24 48 16
run 32 64 16
run 16 32 32
run 24 48 8
run 12 36 16
run 16 32 16 --zerocopy
run 24 72 8 --zerocopy
run 16 64 8
rm -rf /dev/shm/kk_bench_*
echo "== done"
