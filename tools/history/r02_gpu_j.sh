#!/bin/bash
# Round 2, GPU call J (1 GPU): final state — transposing plans on static scheduling (GPT-2 timing + ncu), the full suite, both bench arms.
#   /usr/local/graft/bin/gpurun --timeout 1000 -- 'bash tools/r02/gpu_j.sh'
O=gpurun_out/r02j; mkdir -p $O
echo "== 1. parity + GPT-2 timing"
timeout 90 python tools/gpu_quick.py > $O/quick.stdout 2>&1; echo "rc=$?"; grep -c PASS $O/quick.stdout; grep -v PASS $O/quick.stdout | tail -3 | cut -c1-200
KK_QUICK_OUT=r02j/gpt2_quick.json timeout 120 python tools/gpu_quick_gpt2.py > $O/gpt2.stdout 2>&1; echo "rc=$?"; tail -c 800 $O/gpt2.stdout; echo
echo "== 2. full GPU suite (verbose)"
timeout 700 python -m pytest tests/ -v -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $O/pytest_gpu.log | cut -c1-240; grep -E "FAILED|ERROR" $O/pytest_gpu.log | head
echo "== 3. smoke + bench, both arms (as the driver runs them)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 300 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; head -c 500 $O/bench_ref.json; echo
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "ours rc=$?"; head -c 500 $O/bench_n1.json; echo; tail -2 $O/bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02j/bench_n1.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "time_to_agent_ready_s")}, d["roofline"]["frac"], d["roofline"]["hbm_write_frac"])
    print("secondary", d["secondary"]["ms_per_step"], d["secondary"]["roofline"]["frac"], d["secondary"]["roofline"]["hbm_write_frac"])
    print("secondary_gpt2", d["secondary_gpt2"]["ms_per_step"], d["secondary_gpt2"]["roofline"]["frac"])
    print("e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "probe", d["setup"]["h2d_probe_GBps"], "clocks", d["clocks"])
except Exception as e:
    print("unreadable", e)
PY
echo "== 4. ncu: GPT-2 launch at HEAD, launch list of the kernel-only bench"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:kk_convert_kernel -s 3 -c 1 -o $O/prof_gpt2 -f python tools/gpu_quick_gpt2.py > $O/ncu_gpt2.log 2>&1; echo "ncu rc=$?"
ncu -i $O/prof_gpt2.ncu-rep --page raw --csv > $O/prof_gpt2.raw.csv 2>/dev/null; ncu -i $O/prof_gpt2.ncu-rep --page details > $O/prof_gpt2.details.txt 2>/dev/null
ncu -i $O/prof_gpt2.ncu-rep --page source --csv --print-source sass 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); hdr = rows[1]
ia, ix, iw = hdr.index('Source'), hdr.index('L1 Wavefronts Shared Excessive'), hdr.index('L1 Wavefronts Shared')
out = []
for r in rows[2:]:
    try: x, w = int(r[ix]), int(r[iw])
    except: continue
    if w > 0: out.append((x, w, r[ia].strip()[:50]))
print('shared wavefronts', sum(o[1] for o in out), 'excessive', sum(o[0] for o in out))
for o in sorted(out, reverse=True)[:5]: print('  ', o)
" | tee $O/prof_gpt2_shared_conflicts.txt
rm -f $O/prof_gpt2.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_copy_resident.csv python bench.py --steps 2 --warmup 3 --kernel-only --no-cpu-baseline --no-secondary > $O/bench_under_ncu.log 2>&1; echo "rc=$?"; grep -c kk_convert $O/launches_copy_resident.csv
echo "== done"
