#!/bin/bash
# Round 2, GPU call D (8 GPUs — charged 8x, every step under its own timeout): the BASELINE config 3 numbers in the torchrun shape the driver uses.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 600 -- 'bash tools/r02/gpu_d.sh'
O=gpurun_out/r02d; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
summ() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ("value", "ms_per_step", "time_to_agent_ready_s", "time_to_agent_ready_incl_kk_open_s", "time_to_agent_ready_single_process_s", "time_to_agent_ready_breakdown_rank0", "nccl_compare", "nvls_compare", "pull_stages_ms_rank0")})
    print("  roofline", {k: d["roofline"].get(k) for k in ("bound", "achieved", "peak", "frac", "frac_of_nominal", "stage_ms")}, "e2e", {k: d["e2e"].get(k) for k in ("value", "file_GBps", "ms_per_step")}, "clocks", d.get("clocks"))
except Exception as e:
    print(f, "unreadable:", e)
PY
}
echo "== 1. default bench (PULL order, peers attached while stage 1 runs): time-to-agent-ready, value, nvlink roofline, e2e"
timeout 240 $TR bench.py --gpus 8 --steps 10 --warmup 3 --keep-data > $O/bench_n8_pull.json 2> $O/bench_n8_pull.err; echo "rc=$?"; summ $O/bench_n8_pull.json
echo "== 2. fused P2P stores + ncclAllGather + NVLS multimem.st comparisons (and the one-process-all-GPUs time-to-ready)"
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --fanout p2p --nccl-compare --nvls-compare --keep-data > $O/bench_n8_p2p.json 2> $O/bench_n8_p2p.err; echo "rc=$?"; summ $O/bench_n8_p2p.json
echo "== 3. ncu --set full of one fused launch on device 0 of an 8-GPU single-process context (NVLink byte counters)"
timeout 200 ncu --set full --section Nvlink --section Nvlink_Tables --section Nvlink_Topology --clock-control none --import-source on -k regex:kk_convert_kernel --devices 0 -s 2 -c 1 \
  -o $O/prof_fanout_n8 -f python tools/profile_fanout.py 8 8 > $O/ncu_fanout_n8.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu_fanout_n8.log
ncu -i $O/prof_fanout_n8.ncu-rep --page raw --csv > $O/prof_fanout_n8.raw.csv 2>/dev/null
ncu -i $O/prof_fanout_n8.ncu-rep --page details > $O/prof_fanout_n8.details.txt 2>/dev/null
echo "== 4. multi-GPU pytest file on 8 GPUs"
timeout 240 python -m pytest tests/test_gpu_multi.py tests/test_gpu_quants.py -v -m gpu -p no:cacheprovider -k "test_gpu_multi or pull_one_process or nvls" -rs > $O/pytest_multigpu_n8.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_multigpu_n8.log | cut -c1-200
rm -rf /dev/shm/kk_bench_* /dev/shm/kk_prof_*
echo "== done"
