#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_last.log 2>&1; echo "smoke rc=$?" > gpurun_out/box_last.txt
timeout 200 python -m pytest tests/test_gpu_load.py -m gpu -q -x -k "concurrent or eight_concurrent or virtual_ranks_scatter or mixed_safetensors" > gpurun_out/pytest_last.log 2>&1; echo "pytest rc=$?" >> gpurun_out/box_last.txt
timeout 150 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$?" >> gpurun_out/box_last.txt
cat gpurun_out/box_last.txt; tail -4 gpurun_out/pytest_last.log | cut -c1-200; head -c 400 gpurun_out/bench_last.json; tail -2 gpurun_out/bench_last.err
