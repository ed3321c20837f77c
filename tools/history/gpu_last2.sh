#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_load.py -m gpu -q -x -k "alignment_8 or error_paths or q4_k_m_style or scatter_exchange" > gpurun_out/pytest_last2.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_last2.log | cut -c1-250
