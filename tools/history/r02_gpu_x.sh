#!/bin/bash
# Round 2, GPU call X (1 GPU): the GPU suite once more on HEAD (reader loop touched by the bounce experiment), smoke.
#   /usr/local/graft/bin/gpurun --timeout 200 -- 'bash tools/r02/gpu_x.sh'
O=gpurun_out/r02x; mkdir -p $O
timeout 150 python -m pytest tests/ -q -m gpu -p no:cacheprovider -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log | cut -c1-200
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
echo "== done"
