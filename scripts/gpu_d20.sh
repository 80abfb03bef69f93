#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/perf_upload.py 2>&1 | tee gpurun_out/d20_upload.log
timeout 900 python -m pytest tests -x -q -m gpu -k "not reference and not full_size and not 200k" 2>&1 | tail -2
