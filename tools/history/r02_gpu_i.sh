#!/bin/bash
# Round 2, GPU call I (8 GPUs): the HEAD build (dynamic tile scheduling, rounded IPC buffers, page-cache warm-up) in the shape the driver's scaling run
# uses, on a fresh box: N = 8 default bench, then the N = 2 and N = 8 ncu captures of one fused launch with the NVLink counters, then the multi-GPU tests.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 600 -- 'bash tools/r02/gpu_i.sh'
O=gpurun_out/r02i; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522"
echo "== 1. default bench, N = 8, first run on this box"
timeout 200 $TR bench.py --gpus 8 --steps 10 --warmup 3 --keep-data > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc=$?"
python - $O/bench_n8.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "time_to_agent_ready_s", "time_to_agent_ready_incl_kk_open_s", "time_to_agent_ready_single_process_s", "time_to_agent_ready_single_process_incl_kk_open_s")})
    print("  max  ", d.get("time_to_agent_ready_breakdown_max_over_ranks"))
    print("  roofline", {k: d["roofline"].get(k) for k in ("bound", "achieved", "peak", "frac", "frac_of_nominal", "stage_ms")}, "e2e", {k: d["e2e"].get(k) for k in ("value", "file_GBps", "ms_per_step")})
    print("  setup", d.get("setup"), d.get("clocks"))
except Exception as e:
    print("unreadable:", e)
PY
echo "== 2. ncu --set full, one fused launch, N = 2 then N = 8 (single process, device 0)"
for n in 2 8; do
  timeout 200 ncu --set full --section Nvlink --section Nvlink_Tables --section Nvlink_Topology --clock-control none --import-source on -k regex:kk_convert_kernel --devices 0 -s 2 -c 1 \
    -o $O/prof_fanout_n$n -f python tools/profile_fanout.py $n 8 > $O/ncu_fanout_n$n.log 2>&1; echo "ncu n=$n rc=$?"; tail -2 $O/ncu_fanout_n$n.log
  ncu -i $O/prof_fanout_n$n.ncu-rep --page raw --csv > $O/prof_fanout_n$n.raw.csv 2>/dev/null
  ncu -i $O/prof_fanout_n$n.ncu-rep --page details > $O/prof_fanout_n$n.details.txt 2>/dev/null
  rm -f $O/prof_fanout_n$n.ncu-rep
  grep -E "Transmitted (User )?Bytes|Transmitted Peak|Duration" $O/prof_fanout_n$n.details.txt | head -6
done
echo "== 3. multi-GPU pytest cases"
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_quants.py tests/test_gpu_vmm.py -v -m gpu -p no:cacheprovider -k "test_gpu_multi or pull_one_process or nvls or vmm or read_only" -rs > $O/pytest_multigpu_n8.log 2>&1; echo "rc=$?"; tail -3 $O/pytest_multigpu_n8.log | cut -c1-200
rm -rf /dev/shm/kk_bench_* /dev/shm/kk_prof_*
echo "== done"
