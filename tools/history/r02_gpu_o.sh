#!/bin/bash
# Round 2, GPU call O (1 GPU): how the page-cache bytes reach the pinned slot — pread vs a mapping + memcpy vs non-temporal stores (KUKEON_GPULOAD_READ).
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/r02/gpu_o.sh'
O=gpurun_out/r02o; mkdir -p $O
: > $O/e2e_read_modes.jsonl
run() {  # mode readers [extra flags]
  mode=$1; r=$2; shift 2
  KUKEON_GPULOAD_READ=$mode timeout 120 python bench.py --e2e-only --steps 4 --warmup 2 --readers $r --slots $((2 * r)) --slot-mb 16 --keep-data "$@" 2> $O/e2e_last.err | tail -1 > $O/line.json
  python - "$mode" "$@" <<'PY' >> gpurun_out/r02o/e2e_read_modes.jsonl
import sys, json
try:
    d = json.loads(open("gpurun_out/r02o/line.json").read())
    d["read_mode"] = sys.argv[1]; d["extra"] = sys.argv[2:]
    print(json.dumps(d))
    rd = d.get("readers_last_step") or {}
    sys.stderr.write("%s readers %s %s -> %.2f GB/s %.1f ms; probe %.1f; pread_s %.3f ok=%s\n" % (sys.argv[1], d["config"]["readers"], sys.argv[2:], d["e2e"]["value"], d["e2e"]["ms_per_step"],
                     d.get("h2d_probe_GBps") or 0, (rd.get("pread_s", -1) / max(rd.get("threads", 1), 1)), d["config"].get("verified_vs_files")))
except Exception as e:
    sys.stderr.write("unreadable %s %s\n" % (sys.argv[1], e))
PY
}
run pread 16
run mmap 16
run mmap_nt 16
run populate_nt 16
run mmap_nt 24
run mmap_nt 32
run populate_nt 32
run pread 16 --no-numa-pin
run mmap_nt 16 --no-numa-pin
run mmap_nt 32 --no-numa-pin
tail -5 $O/e2e_last.err | cut -c1-300
echo "== done"
