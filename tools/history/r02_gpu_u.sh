#!/bin/bash
# Round 2, GPU call U (1 GPU): which part of an e2e step carries the occasional 40-150 ms (load_part is steady at 0.30 s): export or checksum.
#   /usr/local/graft/bin/gpurun --timeout 300 -- 'bash tools/r02/gpu_u.sh'
O=gpurun_out/r02u; mkdir -p $O
timeout 200 python bench.py --e2e-only --steps 16 --warmup 2 2> $O/e2e.err | tail -1 > $O/e2e_auto.json
python - <<'PY'
import json
try:
    e = json.loads(open("gpurun_out/r02u/e2e_auto.json").read())
    print(round(e["e2e"]["value"], 2), "GB/s; probe", e["h2d_probe_GBps"])
    for x in e["steps_detail"]:
        print(round(x["ms"]), "load_part", round(x["load_part_ms"]), "export", round(x["export_ms"], 1), "checksum", round(x["checksum_ms"], 1), "close", round(x["files_close_s"] * 1e3, 1))
except Exception as ex:
    print("unreadable", ex)
PY
echo "== done"
