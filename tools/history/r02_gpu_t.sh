#!/bin/bash
# Round 2, GPU call T (2 GPUs): the e2e leg at N = 2, pread against the mapped streaming read path, same box.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash tools/r02/gpu_t.sh'
O=gpurun_out/r02t; mkdir -p $O
run() {
  KUKEON_GPULOAD_READ=$1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --e2e-only --steps 6 --warmup 2 --keep-data 2> $O/e2e_$1.err | tail -1 > $O/e2e_n2_$1.json
  python - $1 <<'PY'
import json, sys
try:
    e = json.loads(open("gpurun_out/r02t/e2e_n2_%s.json" % sys.argv[1]).read())
    print(sys.argv[1], "N=2 e2e", round(e["e2e"]["value"], 1), "GB/s delivered,", round(e["e2e"]["ms_per_step"], 1), "ms/step", [round(x) for x in e["e2e_ms_each"]], "rank0 load_part", [round(x["load_part_ms"]) for x in e["steps_detail"]],
          "rank0 reader copy ms", [round(x["reader_avg"]["pread_s"] * 1e3) for x in e["steps_detail"]], "ttr", round(e["time_to_agent_ready_s"], 3))
except Exception as ex:
    print("unreadable", sys.argv[1], ex)
PY
}
run pread
run auto
tail -3 $O/e2e_auto.err | cut -c1-300
echo "== done"
