#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?" > gpurun_out/box9.txt
cat gpurun_out/box9.txt; tail -15 gpurun_out/pytest_gpu9.log | cut -c1-220
