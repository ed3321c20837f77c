#!/bin/bash
# Round 2, GPU call S (1 GPU): the bench line with the mapped read path and the harness's GC kept out of the e2e steps.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/r02/gpu_s.sh'
O=gpurun_out/r02s; mkdir -p $O
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "ours rc=$?"
timeout 300 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02s/bench_n1.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "time_to_agent_ready_s")}, d["roofline"]["frac"])
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], [round(x) for x in d["e2e"]["ms_each"]], "probe", d["setup"]["h2d_probe_GBps"], "cpu", d["cpu_baseline"]["value"])
    print("ttr breakdown", d["time_to_agent_ready_breakdown_rank0"])
    r = json.loads(open("gpurun_out/r02s/bench_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["ms_per_step"])
except Exception as e:
    print("unreadable", e)
PY
echo "== done"
