#!/bin/bash
# Round 2, GPU call F2 (8 GPUs): the default bench on a FRESH box, twice.  Call F's first run took 1.54 s to agent-ready where the same command as a
# second run takes 0.40 s, and rank 0's own breakdown accounts for 0.19 s of it: some other rank's cold kk_load_part is slow the first time on a box.
# The line now carries the per-component maximum over ranks and where the reader threads spent the cold load (slot wait / pread / issue / drain).
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 400 -- 'bash tools/r02/gpu_f2.sh'
O=gpurun_out/r02f2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
for i in 1 2; do
  timeout 180 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-single-process --keep-data > $O/bench_n8_run$i.json 2> $O/bench_n8_run$i.err; echo "run $i rc=$?"
  python - $O/bench_n8_run$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "time_to_agent_ready_s", "time_to_agent_ready_incl_kk_open_s")})
    print("  rank0", d.get("time_to_agent_ready_breakdown_rank0"))
    print("  max  ", d.get("time_to_agent_ready_breakdown_max_over_ranks"))
    print("  e2e", {k: d["e2e"].get(k) for k in ("value", "file_GBps", "ms_per_step")}, d["roofline"].get("frac"))
except Exception as e:
    print("unreadable:", e)
PY
done
rm -rf /dev/shm/kk_bench_*
echo "== done"
