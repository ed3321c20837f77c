#!/bin/bash
# Second GPU visit: all GPU tests, e2e sweeps, resident-kernel ncu capture, q4_K kernel numbers.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" > gpurun_out/box2.txt
python bench.py --keep-data --e2e-only --steps 3 --warmup 1 > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
for cfg in "4 8 64" "8 16 32" "16 32 32" "16 32 16" "32 64 16" "24 48 32"; do
  set -- $cfg
  python bench.py --keep-data --e2e-only --steps 3 --warmup 1 --readers $1 --slots $2 --slot-mb $3 >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
done
python bench.py --keep-data --e2e-only --steps 3 --warmup 1 --readers 16 --slots 32 --slot-mb 32 --no-numa-pin >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
python bench.py --keep-data --e2e-only --steps 3 --warmup 1 --readers 16 --slots 32 --slot-mb 32 --zerocopy >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
echo "sweep done" >> gpurun_out/box2.txt
# resident kernel: launch list + full capture (every kk_convert launch is a resident per-shard launch)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_resident.csv \
   python bench.py --keep-data --kernel-only --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kk_convert -s 6 -c 2 -o gpurun_out/prof_copy_resident \
   python bench.py --keep-data --kernel-only --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
echo "ncu copy done" >> gpurun_out/box2.txt
rm -rf /dev/shm/kk_bench_llama*
# q4_K: Mixtral reduced to 4 layers (5.8 GB in -> 20.7 GB out)
python bench.py --workload mixtral-q4k --layers 4 --keep-data --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q4k_l4.json 2> gpurun_out/bench_q4k_l4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kk_convert -s 2 -c 1 -o gpurun_out/prof_q4k_resident \
   python bench.py --workload mixtral-q4k --layers 4 --keep-data --kernel-only --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1
echo "q4k done" >> gpurun_out/box2.txt
python bench.py --workload gpt2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gpt2.json 2> gpurun_out/bench_gpt2.err
cat gpurun_out/box2.txt; tail -5 gpurun_out/pytest_gpu2.log; cat gpurun_out/sweep.jsonl; tail -3 gpurun_out/sweep.err; cat gpurun_out/bench_q4k_l4.json; tail -3 gpurun_out/bench_q4k_l4.err
