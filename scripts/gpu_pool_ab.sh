#!/bin/bash
# pool-mode cast kernel: correctness first (small tests), then perf A/B of the variants given as arguments
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cast or cube or find_closest or golden or shard or degenerate or random_transformed" > gpurun_out/pool_pytest_small.log 2>&1
rc=$?; echo "small pytest rc=$rc"; tail -5 gpurun_out/pool_pytest_small.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/pool_pytest_small.log; fi
for spec in "$@"; do
  v=${spec%%:*}; budget=${spec#*:}
  if [ "$v" = default ]; then unset J3DG_LIB; else export J3DG_LIB=build/variants/libj3dg_$v.so; fi
  if [ "$budget" != "$spec" ]; then export J3DG_LANE_BUDGET=$budget; else unset J3DG_LANE_BUDGET; fi
  echo -n "[$spec] "
  timeout 200 python scripts/perf_cast.py 2>&1 | grep -E "^lib=|timeline|rror" | tail -3
done | tee gpurun_out/pool_ab.log
if [ $rc -eq 0 ]; then
  unset J3DG_LIB J3DG_LANE_BUDGET
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pool_pytest_all.log 2>&1; echo "full pytest rc=$?"; tail -5 gpurun_out/pool_pytest_all.log
fi
