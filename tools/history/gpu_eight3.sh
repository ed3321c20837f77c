#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8c.json 2> gpurun_out/bench_n8c.err; echo "n8 rc=$?" > gpurun_out/box8c.txt
$TR --nproc-per-node 8 --master-port 29552 bench.py --gpus 8 --workload llama3-70b-scatter --steps 3 --warmup 2 > gpurun_out/bench_scatter_n8c.json 2> gpurun_out/bench_scatter_n8c.err; echo "scatter n8 rc=$?" >> gpurun_out/box8c.txt
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "scatter_exchange or raw" > gpurun_out/pytest_gpu8c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/box8c.txt
rm -rf /dev/shm/kk_bench_*
cat gpurun_out/box8c.txt; tail -3 gpurun_out/pytest_gpu8c.log | cut -c1-200; tail -3 gpurun_out/bench_scatter_n8c.err | cut -c1-300
