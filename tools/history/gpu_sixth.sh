#!/bin/bash
# 2 GPUs: full GPU test suite (incl. RAW fan-out, odd transposes, full-size properties), q4_K broadcast order A/B at N=2.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?" > gpurun_out/box6.txt
$TR --nproc-per-node 2 --master-port 29531 bench.py --gpus 2 --workload mixtral-q4k --layers 8 --steps 3 --warmup 3 --keep-data > gpurun_out/mix8_p2p_n2.json 2> gpurun_out/mix8_p2p_n2.err; echo "p2p rc=$?" >> gpurun_out/box6.txt
$TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --workload mixtral-q4k --layers 8 --steps 3 --warmup 3 --fanout raw > gpurun_out/mix8_raw_n2.json 2> gpurun_out/mix8_raw_n2.err; echo "raw rc=$?" >> gpurun_out/box6.txt
$TR --nproc-per-node 2 --master-port 29533 bench.py --gpus 2 --workload llama3-70b-scatter --layers 8 --steps 3 --warmup 3 > gpurun_out/scatter8_n2.json 2> gpurun_out/scatter8_n2.err; echo "scatter rc=$?" >> gpurun_out/box6.txt
cat gpurun_out/box6.txt; tail -12 gpurun_out/pytest_gpu6.log | cut -c1-220
