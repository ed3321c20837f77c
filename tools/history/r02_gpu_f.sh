#!/bin/bash
# Round 2, GPU call F (8 GPUs, charged 8x): which fan-out order is the default of the one-process-per-GPU shape?  Call D measured the P2P-store order
# at 0.40 s time-to-ready (seven 16 GB pool mappings in 0.07 s — 3.6 s in round 1, before pools were rounded to 2 MiB multiples) but as the SECOND run
# on that box; here it runs FIRST on a fresh box, then PULL with its slice buffers rounded the same way, then two reader-count points of the e2e leg.
# Files are written with their page-cache pages interleaved over the NUMA nodes (bench.py make_files).
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 600 -- 'bash tools/r02/gpu_f.sh'
O=gpurun_out/r02f; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
summ() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ("value", "ms_per_step", "time_to_agent_ready_s", "time_to_agent_ready_incl_kk_open_s", "time_to_agent_ready_breakdown_rank0", "pull_stages_ms_rank0")})
    if "roofline" in d:
        print("  roofline", {k: d["roofline"].get(k) for k in ("bound", "achieved", "peak", "frac", "frac_of_nominal", "stage_ms")})
    print("  e2e", {k: d["e2e"].get(k) for k in ("value", "file_GBps", "ms_per_step")}, "files:", (d.get("config") or {}).get("files"))
except Exception as e:
    print(f, "unreadable:", e)
PY
}
echo "== 1. P2P stores, first run on this box"
timeout 200 $TR bench.py --gpus 8 --steps 10 --warmup 3 --fanout p2p --no-single-process --keep-data > $O/bench_n8_p2p_first.json 2> $O/bench_n8_p2p_first.err; echo "rc=$?"; summ $O/bench_n8_p2p_first.json
echo "== 2. PULL, slice buffers in 2 MiB multiples"
timeout 200 $TR bench.py --gpus 8 --steps 10 --warmup 3 --fanout pull --no-single-process --keep-data > $O/bench_n8_pull2.json 2> $O/bench_n8_pull2.err; echo "rc=$?"; summ $O/bench_n8_pull2.json
echo "== 3. e2e leg, P2P, reader threads per rank 8 and 32 (default 16)"
for R in 8 32; do
  timeout 120 $TR bench.py --gpus 8 --steps 6 --warmup 3 --fanout p2p --e2e-only --readers $R --no-single-process --keep-data > $O/e2e_n8_readers$R.json 2> $O/e2e_n8_readers$R.err; echo "readers $R rc=$?"; cat $O/e2e_n8_readers$R.json | cut -c1-600
done
echo "== 4. N = 4, P2P (profiles only; the driver's scaling run measures it too)"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 4 --steps 10 --warmup 3 --fanout p2p --no-single-process --keep-data > $O/bench_n4_p2p.json 2> $O/bench_n4_p2p.err; echo "rc=$?"; summ $O/bench_n4_p2p.json
rm -rf /dev/shm/kk_bench_*
echo "== done"
