#!/bin/bash
# Third GPU visit (2 GPUs): multi-GPU tests, q4_K kernel A/B (8 vs 16 consumer warps), N=2 bench, fabric probes.
mkdir -p gpurun_out
python - > gpurun_out/probe.txt 2>&1 <<'PY'
from cuda.bindings import driver as cu, runtime as rt
print(cu.cuInit(0))
err, n = cu.cuDeviceGetCount(); print("devices", n)
for d in range(n):
    err, dev = cu.cuDeviceGet(d)
    for name in ("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED",
                 "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED",
                 "CU_DEVICE_ATTRIBUTE_GPU_DIRECT_RDMA_SUPPORTED"):
        a = getattr(cu.CUdevice_attribute, name, None)
        if a is not None:
            print(d, name, cu.cuDeviceGetAttribute(a, dev))
for i in range(n):
    for j in range(n):
        if i != j:
            print("canAccessPeer", i, j, rt.cudaDeviceCanAccessPeer(i, j))
PY
nvidia-smi topo -m >> gpurun_out/probe.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" > gpurun_out/box3.txt
python bench.py --workload mixtral-q4k --layers 4 --keep-data --kernel-only --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q4k_w8.json 2> gpurun_out/q4k_w8.err
KUKEON_GPULOAD_LIB=$PWD/kukeon_b200/csrc/build/libkk_w16.so python bench.py --workload mixtral-q4k --layers 4 --keep-data --kernel-only --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q4k_w16.json 2> gpurun_out/q4k_w16.err
KUKEON_GPULOAD_LIB=$PWD/kukeon_b200/csrc/build/libkk_w16.so python bench.py --workload gpt2 --kernel-only --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/gpt2_w16.json 2> gpurun_out/gpt2_w16.err
python bench.py --workload gpt2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/gpt2_w8.json 2> gpurun_out/gpt2_w8.err
rm -rf /dev/shm/kk_bench_mixtral*
echo "q4k ab done" >> gpurun_out/box3.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --nccl-compare --keep-data > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?" >> gpurun_out/box3.txt
python bench.py --steps 5 --warmup 3 --keep-data > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; echo "bench n1 rc=$?" >> gpurun_out/box3.txt
cat gpurun_out/box3.txt; tail -8 gpurun_out/pytest_gpu3.log; head -c 600 gpurun_out/probe.txt
