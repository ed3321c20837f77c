#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?" > gpurun_out/box7.txt
python bench.py --workload mixtral-q4k --layers 4 --keep-data --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q4k_final.json 2> gpurun_out/q4k_final.err; echo "q4k rc=$?" >> gpurun_out/box7.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kk_convert -s 2 -c 1 -o gpurun_out/prof_q4k_v2 \
   python bench.py --workload mixtral-q4k --layers 4 --keep-data --kernel-only --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_q4k_v2.log 2>&1
rm -rf /dev/shm/kk_bench_mixtral*
python bench.py --workload gpt2 --keep-data --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/gpt2_final.json 2> gpurun_out/gpt2_final.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kk_convert -s 2 -c 1 -o gpurun_out/prof_gpt2 \
   python bench.py --workload gpt2 --keep-data --kernel-only --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gpt2.log 2>&1
rm -rf /dev/shm/kk_bench_gpt2*
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo "n1 rc=$?" >> gpurun_out/box7.txt
cat gpurun_out/box7.txt; tail -8 gpurun_out/pytest_gpu7.log | cut -c1-200
