#!/bin/bash
# Second 8-GPU visit: multi-GPU parity (incl. RAW order), N=8 broadcast with eager peer setup (time-to-ready),
# Mixtral q4_K broadcast in both orders at full size, Llama-3-70B scatter with mmap row gathers.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_gpu8b.log 2>&1; echo "pytest multi rc=$?" > gpurun_out/box8b.txt
$TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 --nccl-compare --keep-data > gpurun_out/bench_n8b.json 2> gpurun_out/bench_n8b.err; echo "n8 rc=$?" >> gpurun_out/box8b.txt
$TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 3 --warmup 3 --lazy-peers > gpurun_out/bench_n8_lazy.json 2> gpurun_out/bench_n8_lazy.err; echo "n8 lazy rc=$?" >> gpurun_out/box8b.txt
rm -rf /dev/shm/kk_bench_llama3-8b*
$TR --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --workload mixtral-q4k --steps 3 --warmup 3 --fanout raw --keep-data > gpurun_out/bench_mixtral_raw_n8.json 2> gpurun_out/bench_mixtral_raw_n8.err; echo "mixtral raw n8 rc=$?" >> gpurun_out/box8b.txt
$TR --nproc-per-node 8 --master-port 29544 bench.py --gpus 8 --workload mixtral-q4k --steps 3 --warmup 3 > gpurun_out/bench_mixtral_p2p_n8.json 2> gpurun_out/bench_mixtral_p2p_n8.err; echo "mixtral p2p n8 rc=$?" >> gpurun_out/box8b.txt
rm -rf /dev/shm/kk_bench_mixtral*
$TR --nproc-per-node 8 --master-port 29545 bench.py --gpus 8 --workload llama3-70b-scatter --steps 3 --warmup 2 > gpurun_out/bench_scatter_n8b.json 2> gpurun_out/bench_scatter_n8b.err; echo "scatter n8 rc=$?" >> gpurun_out/box8b.txt
rm -rf /dev/shm/kk_bench_*
cat gpurun_out/box8b.txt; tail -4 gpurun_out/pytest_gpu8b.log | cut -c1-200
