#!/bin/bash
# Round 2, GPU call R (1 GPU): the mapped streaming read path as the default for tmpfs shards; checksum without cudaMalloc / cudaFree per call.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r02/gpu_r.sh'
O=gpurun_out/r02r; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_load.py -q -m gpu -p no:cacheprovider -k "tmpfs_shards or medium_checkpoint or mixed_safetensors" > $O/pytest_new.log 2>&1; echo "new tests rc=$?"; tail -2 $O/pytest_new.log | cut -c1-200
timeout 700 python -m pytest tests/ -q -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "ours rc=$?"
KUKEON_GPULOAD_READ=pread timeout 200 python bench.py --e2e-only --steps 6 --warmup 2 2> $O/e2e_pread.err | tail -1 > $O/e2e_pread.json
timeout 200 python bench.py --e2e-only --steps 6 --warmup 2 2> $O/e2e_auto.err | tail -1 > $O/e2e_auto.json
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02r/bench_n1.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "time_to_agent_ready_s")}, d["roofline"]["frac"])
    print("secondary", d["secondary"]["roofline"]["frac"], d["secondary"]["roofline"]["hbm_write_frac"], "gpt2", d["secondary_gpt2"]["ms_per_step"], d["secondary_gpt2"]["roofline"]["frac"])
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "probe", d["setup"]["h2d_probe_GBps"], "cpu", d["cpu_baseline"]["value"])
    r = json.loads(open("gpurun_out/r02r/bench_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["ms_per_step"])
    for n in ("pread", "auto"):
        e = json.loads(open("gpurun_out/r02r/e2e_%s.json" % n).read())
        print(n, round(e["e2e"]["value"], 2), [round(x) for x in e["e2e_ms_each"]], [round(x["load_part_ms"]) for x in e["steps_detail"]])
except Exception as e:
    print("unreadable", e)
PY
echo "== done"
