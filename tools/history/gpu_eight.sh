#!/bin/bash
# 8-GPU visit: multi-GPU parity tests, N=8/4/2 broadcast bench (fused P2P fan-out vs NCCL all-gather), Mixtral q4_K
# broadcast and Llama-3-70B scatter at full size, reference arm.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest multi rc=$?" > gpurun_out/box8.txt
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --nccl-compare --keep-data > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "n8 rc=$?" >> gpurun_out/box8.txt
$TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 5 --warmup 3 --nccl-compare --keep-data > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "n4 rc=$?" >> gpurun_out/box8.txt
$TR --nproc-per-node 2 --master-port 29523 bench.py --gpus 2 --steps 5 --warmup 3 --nccl-compare --keep-data > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?" >> gpurun_out/box8.txt
python bench.py --gpus 1 --steps 5 --warmup 3 --keep-data > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 rc=$?" >> gpurun_out/box8.txt
$TR --nproc-per-node 8 --master-port 29524 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; echo "ref rc=$?" >> gpurun_out/box8.txt
rm -rf /dev/shm/kk_bench_llama3-8b*
$TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --workload mixtral-q4k --steps 3 --warmup 3 > gpurun_out/bench_mixtral_n8.json 2> gpurun_out/bench_mixtral_n8.err; echo "mixtral n8 rc=$?" >> gpurun_out/box8.txt
rm -rf /dev/shm/kk_bench_mixtral*
$TR --nproc-per-node 8 --master-port 29526 bench.py --gpus 8 --workload llama3-70b-scatter --steps 3 --warmup 3 > gpurun_out/bench_scatter_n8.json 2> gpurun_out/bench_scatter_n8.err; echo "scatter n8 rc=$?" >> gpurun_out/box8.txt
rm -rf /dev/shm/kk_bench_*
cat gpurun_out/box8.txt; tail -4 gpurun_out/pytest_gpu8.log; for f in n8 n4 n2 n1 mixtral_n8 scatter_n8; do tail -2 gpurun_out/bench_$f.err; done
