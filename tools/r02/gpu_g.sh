#!/bin/bash
# Round 2, GPU call G (1 GPU): e2e leg tuning at N = 1 (>= 0.93 of the pinned-H2D probe; 0.86 at the defaults 16 readers x 32 slots x 16 MiB).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r02/gpu_g.sh'
O=gpurun_out/r02g; mkdir -p $O
: > $O/e2e_sweep.jsonl
run() {  # readers slots slot_mb [extra flags]
  r=$1; s=$2; mb=$3; shift 3
  timeout 150 python bench.py --e2e-only --steps 4 --warmup 2 --readers $r --slots $s --slot-mb $mb --keep-data "$@" 2> $O/e2e_last.err | tail -1 >> $O/e2e_sweep.jsonl
  tail -1 $O/e2e_sweep.jsonl | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('readers', d['config']['readers'], 'slots', d['config']['slots'], 'slot_mb', d['config']['slot_mb'], 'zerocopy', d['config']['zerocopy'], '->', round(d['e2e']['value'], 2), 'GB/s', round(d['e2e']['ms_per_step'], 1), 'ms; probe', round(d['h2d_probe_GBps'] or 0, 1))
except Exception as e:
    print('unreadable', e)
"
}
run 16 32 16
run 16 48 16
run 24 48 16
run 32 64 16
run 16 32 32
run 24 48 8
run 12 36 16
run 16 32 16 --zerocopy
run 24 72 8 --zerocopy
run 16 64 8
rm -rf /dev/shm/kk_bench_*
echo "== done"
