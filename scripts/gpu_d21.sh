#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/d21_knn_launches.csv python scripts/perf_knn.py 2000000 10 > gpurun_out/d21_knn.log 2>&1
timeout 300 python scripts/perf_knn.py 2000000 10 2>&1 | tail -1
timeout 300 python scripts/perf_knn.py 10000000 16 2>&1 | tail -1
python - <<'PY'
import csv,collections
lines=[l for l in open('gpurun_out/d21_knn_launches.csv') if not l.startswith('==')]
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for row in csv.DictReader(lines):
    try: v=float(row['Metric Value'].replace(',',''))
    except: continue
    agg[row['Kernel Name'][:50]][row['Metric Name']].append(v)
for k,m in agg.items():
    t=m['gpu__time_duration.sum']; print(f"{k:50s} n={len(t):3d} total_us={sum(t)/1e3:9.1f} mean_us={sum(t)/len(t)/1e3:9.1f} rdMB={sum(m['dram__bytes_read.sum'])/len(t)/1e6:8.1f} wrMB={sum(m['dram__bytes_write.sum'])/len(t)/1e6:8.1f}")
PY
