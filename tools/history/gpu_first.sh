#!/bin/bash
# First GPU visit: probe the box, run the GPU tests, a reduced bench, the launch list and one full ncu capture.
mkdir -p gpurun_out
{
echo "== box"; nproc; free -g | head -2; df -h /dev/shm /tmp | cat; lscpu | grep -E "Model name|Socket|Thread|NUMA node\(s\)"; nvidia-smi -L
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv
echo "== tmpfs first-touch write"; dd if=/dev/zero of=/dev/shm/ddtest bs=8M count=128 2>&1 | tail -1
echo "== tmpfs re-read"; dd if=/dev/shm/ddtest of=/dev/null bs=8M 2>&1 | tail -1; rm -f /dev/shm/ddtest
} > gpurun_out/box.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/box.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/box.txt
timeout 600 python bench.py --layers 4 --steps 3 --warmup 3 --keep-data > gpurun_out/bench_l4.json 2> gpurun_out/bench_l4.err; echo "bench_l4 rc=$?" >> gpurun_out/box.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_l4.csv \
   python bench.py --layers 4 --steps 2 --warmup 3 --keep-data --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/box.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:kk_convert -s 8 -c 2 -o gpurun_out/prof_copy_l4 \
   python bench.py --layers 4 --steps 1 --warmup 3 --keep-data --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/box.txt
rm -rf /dev/shm/kk_bench_*
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full rc=$?" >> gpurun_out/box.txt
cat gpurun_out/box.txt; tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_l4.json; tail -3 gpurun_out/bench_l4.err; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
