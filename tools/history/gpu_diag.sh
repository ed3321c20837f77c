#!/bin/bash
mkdir -p gpurun_out
python tools/diag_transpose.py > gpurun_out/diag_w16.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_load.py -m gpu -q -k "gpt2 or multi_destination or special or mixed" > gpurun_out/pytest_diag.log 2>&1; echo "pytest rc=$?"
python bench.py --workload gpt2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/gpt2_tma.json 2> gpurun_out/gpt2_tma.err
head -20 gpurun_out/diag_w16.txt; tail -8 gpurun_out/pytest_diag.log; python -c "
import json; d=json.load(open('gpurun_out/gpt2_tma.json')); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
