#!/bin/bash
# Round 2, GPU call C (2 GPUs): everything that needs more than one device, at the cheapest N — the multi-GPU pytest cases (test_gpu_multi.py, PULL over
# CUDA IPC, NVLS), the torchrun bench in both fan-out orders with the NCCL and NVLS comparisons, and one ncu --set full capture of a fused N = 2 launch
# with the NVLink byte counters.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/r02/gpu_c.sh'
O=gpurun_out/r02c; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
echo "== 1. multi-GPU pytest cases"
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_quants.py tests/test_gpu_vmm.py -q -m gpu -p no:cacheprovider -k "test_gpu_multi or pull_one_process or nvls or vmm or read_only or ipc_mount" -rs > $O/pytest_multigpu.log 2>&1; echo "rc=$?"; tail -12 $O/pytest_multigpu.log | cut -c1-240
echo "== 2. torchrun bench, default fan-out (PULL, peers attached while stage 1 runs)"
timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 --keep-data > $O/bench_n2_pull.json 2> $O/bench_n2_pull.err; echo "rc=$?"; head -c 2500 $O/bench_n2_pull.json; echo
echo "== 3. torchrun bench, fused P2P stores + NCCL all-gather + NVLS comparisons"
timeout 400 $TR bench.py --gpus 2 --steps 5 --warmup 3 --fanout p2p --nccl-compare --nvls-compare --keep-data > $O/bench_n2_p2p.json 2> $O/bench_n2_p2p.err; echo "rc=$?"; head -c 2500 $O/bench_n2_p2p.json; echo
python - <<'PY'
import json
for f in ("bench_n2_pull", "bench_n2_p2p"):
    try:
        d = json.loads(open(f"gpurun_out/r02c/{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "time_to_agent_ready_s", "time_to_agent_ready_breakdown_rank0", "nccl_compare", "nvls_compare", "pull_stages_ms_rank0")})
        print("  roofline", {k: d["roofline"].get(k) for k in ("bound", "achieved", "peak", "frac", "frac_of_nominal", "stage_ms")}, "e2e", d["e2e"].get("value"), d["e2e"].get("file_GBps"), d["e2e"].get("ms_per_step"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
echo "== 4. ncu --set full of one fused convert + fan-out launch at N = 2 (one process owning both GPUs; never a multi-rank command)"
timeout 300 ncu --set full --section Nvlink --section Nvlink_Tables --section Nvlink_Topology --clock-control none --import-source on -k regex:kk_convert_kernel --devices 0 -s 2 -c 1 \
  -o $O/prof_fanout_n2 -f python tools/profile_fanout.py 2 8 > $O/ncu_fanout_n2.log 2>&1; echo "ncu rc=$?"; tail -4 $O/ncu_fanout_n2.log
ncu -i $O/prof_fanout_n2.ncu-rep --page raw --csv > $O/prof_fanout_n2.raw.csv 2>/dev/null
ncu -i $O/prof_fanout_n2.ncu-rep --page details > $O/prof_fanout_n2.details.txt 2>/dev/null
python - <<'PY'
import csv
try:
    rows = list(csv.reader(open("gpurun_out/r02c/prof_fanout_n2.raw.csv")))
    hdr, unit, val = rows[0], rows[1], rows[2]
    for i, h in enumerate(hdr):
        if any(k in h for k in ("nvltx__bytes", "nvlrx__bytes", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "lts__t_sectors_srcunit_ltcfabric", "pcie__")) and not h.endswith(("pct", "per_second")):
            print(h, val[i], unit[i])
except Exception as e:
    print("no raw page:", e)
PY
rm -rf /dev/shm/kk_bench_* /dev/shm/kk_prof_*
echo "== done"
