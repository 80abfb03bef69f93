#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not reference and not full_size" 2>&1 | tail -3 | tee gpurun_out/d6_pytest.log
( LANES=2 timeout 200 python scripts/perf_overlap.py 1184 240 2>&1 | tail -1
  LANES=3 timeout 200 python scripts/perf_overlap.py 1184 120 2>&1 | tail -1
  LANES=4 timeout 200 python scripts/perf_overlap.py 1184 120 2>&1 | tail -1 ) | tee gpurun_out/d6_overlap.log
