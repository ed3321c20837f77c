#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" > gpurun_out/box5.txt
python bench.py --workload gpt2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/gpt2_tma.json 2> gpurun_out/gpt2_tma.err; echo "gpt2 rc=$?" >> gpurun_out/box5.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err; echo "n1 rc=$?" >> gpurun_out/box5.txt
cat gpurun_out/box5.txt; tail -15 gpurun_out/pytest_gpu5.log
