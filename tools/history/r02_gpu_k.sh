#!/bin/bash
# Round 2, GPU call K (2 GPUs): same-box A/B of the two tile schedulers (KUKEON_GPULOAD_SCHED), the parallel per-device kk_open, N = 2 bench at HEAD.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash tools/r02/gpu_k.sh'
O=gpurun_out/r02k; mkdir -p $O
echo "== 1. scheduler A/B on one box: GPT-2 (transposes) and 1 Gi-weight BF16 / Q4_K / Q8_0, three processes each"
for s in static dynamic static dynamic static dynamic; do
  KUKEON_GPULOAD_SCHED=$s KK_QUICK_OUT=r02k/gpt2_$s.json timeout 90 python tools/gpu_quick_gpt2.py 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$s', 'gpt2 transposed', round(d['transposed_bf16']['ms_median'], 4), 'plain', round(d['plain_cast_bf16']['ms_median'], 4), 'keep_f32', round(d['transposed_keep_f32']['ms_median'], 4))"
done
for s in static dynamic; do
  KUKEON_GPULOAD_SCHED=$s KK_QUICK_OUT=r02k/types_$s.json timeout 200 python tools/gpu_quick_types.py --types BF16,Q4_K,Q8_0,Q6_K,IQ2_S --weights-m 1024 --passes 10 2>/dev/null | grep -E "^(BF16|Q4_K|Q8_0|Q6_K|IQ2_S) " | python -c "
import sys, json
for l in sys.stdin:
    k, j = l.split(' ', 1); d = json.loads(j); print('$s', k, round(d['ms_median'], 4), round(d['frac_of_copy_peak'], 3))"
done
echo "== 2. torchrun bench, N = 2 (default), with the one-process-all-GPUs kk_open timing"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02k/bench_n2.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "time_to_agent_ready_s", "time_to_agent_ready_single_process_s", "time_to_agent_ready_single_process_incl_kk_open_s", "single_process")})
    print("  roofline", {k: d["roofline"].get(k) for k in ("bound", "achieved", "peak", "frac")}, "e2e", {k: d["e2e"].get(k) for k in ("value", "file_GBps", "ms_per_step")})
except Exception as e:
    print("unreadable", e)
PY
echo "== 3. multi-GPU + core tests on the parallel kk_open build"
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_vmm.py tests/test_zz_gpu_errors.py -q -m gpu -p no:cacheprovider > $O/pytest.log 2>&1; echo "rc=$?"; tail -3 $O/pytest.log | cut -c1-200
rm -rf /dev/shm/kk_bench_*
echo "== done"
