#!/bin/bash
mkdir -p gpurun_out
for c in 0 16 37 74 148; do
  J3DG_CONSUMER_BLOCKS=$c timeout 300 python scripts/perf_overlap.py 1184 200 2>&1 | tail -2
done | tee gpurun_out/d2_overlap.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not reference and not cpp" 2>&1 | tail -3 | tee gpurun_out/d2_pytest.log
