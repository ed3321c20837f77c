#!/bin/bash
# Round 2, GPU call Y (1 GPU, < 1 min): HEAD after the read-path clean-up — the loads that go through the mapped and the pread path.
#   /usr/local/graft/bin/gpurun --timeout 60 -- 'bash tools/r02/gpu_y.sh'
O=gpurun_out/r02y; mkdir -p $O
timeout 55 python -m pytest tests/test_gpu_load.py -q -m gpu -p no:cacheprovider -x -k "tmpfs_shards or mixed_safetensors or medium_checkpoint or llama_multishard or q4_k_m_style" > $O/pytest_subset.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_subset.log | cut -c1-200
echo "== done"
