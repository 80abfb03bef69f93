"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump (first kernel)."""
import csv, sys
path = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
rows = list(csv.reader(open(path)))
end = next((i for i, r in enumerate(rows) if i > 3 and r and r[0] == "File Path"), len(rows))
rows = rows[:end]
hdr = rows[2]; n = len(hdr)
pos = {}
for i, k in enumerate(hdr):
    pos.setdefault(k, i - n)   # negative index from the end: robust against commas in the source column
S, IE, TE, LSB, WAIT, BAR = (pos[k] for k in ("# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_long_sb", "stall_wait", "stall_barrier"))
lines = []
for r in rows[3:]:
    if r and r[0].isdigit() and len(r) >= n:
        src = ",".join(r[1:len(r) - n + 2])
        lines.append((int(r[0]), src, int(r[S]), int(r[IE]), int(r[TE]), int(r[LSB]), int(r[WAIT]), int(r[BAR])))
tot = sum(l[2] for l in lines); ti = sum(l[3] for l in lines); tt = sum(l[4] for l in lines)
print(f"samples {tot}  warp-inst {ti}  thread-inst {tt}  avg threads/inst {tt / ti:.1f}")
for l in sorted(lines):
    if l[2] > tot * thr or l[3] > ti * thr:
        print(f"{l[0]:5d} smp {100 * l[2] / tot:5.1f}% inst {100 * l[3] / ti:5.1f}% thr {l[4] / max(l[3], 1):5.1f} long_sb {l[5]:6d} wait {l[6]:5d} bar {l[7]:5d} | {l[1][:100]}")
