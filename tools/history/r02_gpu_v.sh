#!/bin/bash
# Round 2, GPU call V (1 GPU): final validation of HEAD — full GPU suite, smoke, both bench arms.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r02/gpu_v.sh'
O=gpurun_out/r02v; mkdir -p $O
timeout 700 python -m pytest tests/ -q -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "ours rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02v/bench_n1.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "time_to_agent_ready_s")}, d["roofline"]["frac"])
    print("secondary", d["secondary"]["roofline"]["frac"], d["secondary"]["roofline"]["hbm_write_frac"], "gpt2", d["secondary_gpt2"]["ms_per_step"], d["secondary_gpt2"]["roofline"]["frac"])
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], [round(x) for x in d["e2e"]["ms_each"]], "probe", d["setup"]["h2d_probe_GBps"], "cpu", d["cpu_baseline"]["value"])
    print([(round(x["load_part_ms"]), round(x["export_ms"], 1), round(x["checksum_ms"], 1)) for x in d["e2e"]["steps_detail_rank0"]])
    r = json.loads(open("gpurun_out/r02v/bench_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["ms_per_step"])
except Exception as e:
    print("unreadable", e)
PY
echo "== done"
