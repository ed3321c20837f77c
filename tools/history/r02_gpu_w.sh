#!/bin/bash
# Round 2, GPU call W (2 GPUs, one process each): e2e leg with the default read path against the cache-resident bounce ring (KUKEON_GPULOAD_READ=bounce);
# parity of the bounce path on the loads that mix long and short ranges.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 300 -- 'bash tools/r02/gpu_w.sh'
O=gpurun_out/r02w; mkdir -p $O
run() {
  KUKEON_GPULOAD_READ=$1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --e2e-only --steps 6 --warmup 2 --keep-data 2> $O/e2e_$1.err | tail -1 > $O/e2e_n2_$1.json
  python - $1 <<'PY'
import json, sys
try:
    e = json.loads(open("gpurun_out/r02w/e2e_n2_%s.json" % sys.argv[1]).read())
    print(sys.argv[1], "N=2 e2e", round(e["e2e"]["value"], 1), "GB/s delivered,", round(e["e2e"]["ms_per_step"], 1), "ms/step", [round(x) for x in e["e2e_ms_each"]], "rank0 load_part", [round(x["load_part_ms"]) for x in e["steps_detail"]],
          "copy ms", [round(x["reader_avg"]["pread_s"] * 1e3) for x in e["steps_detail"]], "wait ms", [round(x["reader_avg"]["slot_wait_s"] * 1e3) for x in e["steps_detail"]], "ttr", round(e["time_to_agent_ready_s"], 3), "verified", e.get("verified_vs_files"))
except Exception as ex:
    print("unreadable", sys.argv[1], ex)
PY
}
run auto
run bounce
KUKEON_GPULOAD_READ=bounce timeout 100 python -m pytest tests/test_gpu_load.py -q -m gpu -p no:cacheprovider -k "tmpfs_shards or mixed_safetensors or llama_multishard or virtual_ranks_scatter or q4_k_m_style or gpt2_conv1d" > $O/pytest_bounce.log 2>&1; echo "bounce parity rc=$?"; tail -2 $O/pytest_bounce.log | cut -c1-200
tail -3 $O/e2e_bounce.err | cut -c1-300
echo "== done"
