#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "many_objects" > gpurun_out/d14_top.log 2>&1; echo "top rc=$?"; grep -E "objects @|passed|failed|Error|assert" gpurun_out/d14_top.log | head -20
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not reference and not full_size and not many_objects" 2>&1 | tail -2
