"""launch_table.py LAUNCHES.csv — per-kernel launches / mean time / DRAM bytes of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log."""
import csv, collections, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith('=='))]
h = rows[0]; ki = h.index('Kernel Name'); mi = h.index('Metric Name'); vi = h.index('Metric Value')
d = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi: continue
    name = re.sub(r'\(.*', '', r[ki]).split('::')[-1]
    d.setdefault(name, collections.defaultdict(list))[r[mi]].append(float(r[vi].replace(',', '')))
for k, m in d.items():
    t = m['gpu__time_duration.sum']; n = len(t)
    print(f"{k:28s} n={n:4d} mean_us={sum(t)/n/1e3:9.1f} total_us={sum(t)/1e3:10.1f} rd_MB={sum(m['dram__bytes_read.sum'])/n/1e6:8.1f} wr_MB={sum(m['dram__bytes_write.sum'])/n/1e6:8.1f}")
