"""ncu_regions.py SOURCE.csv — exact warp-instruction accounting of a `ncu --page source --csv --print-source cuda,sass`
dump of cast_kernel: every SASS address is counted ONCE and attributed to the innermost call site inside cast.cu
(the largest line number it appears under, below the kernel wrapper), then summed per named region of cast.cu."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[2]; n = len(hdr)
pos = {k: i - n for i, k in enumerate(hdr)}
end = next((i for i, r in enumerate(rows) if i > 3 and r and r[0] == "File Path"), len(rows))
src = open("j3d_b200/csrc/cast.cu").read().splitlines()
def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]:
            return i + 1
    return None
kernel_line = find("cast_kernel(const TraceParams p)")
marks = [("helpers (tile grid, world_ray)", 1), ("group_loop", find("__device__ __forceinline__ void group_loop")), ("group: queue claim / poll", find("groups without a ray claim the next queue entry")),
         ("group: pool refill (unused here)", find("groups without a ray take the next slots")), ("group: node step", find("(B) inner node: lane c tests child c")), ("group: leaf step", find("(C) leaf: lane c tests triangle c")),
         ("group_kernel / lane_ray_setup", find("group_kernel(const TraceParams p)")), ("lane_loop", find("__device__ __forceinline__ void lane_loop")),
         ("pool: prologue + lambdas", find("__device__ __forceinline__ void pool_loop")), ("pool: classify", find("classify the 64 slots")), ("pool: refill / tile set-up", find("refill free slots from the parked tile")),
         ("pool: select + compact", find("a node step serves 32 rays")), ("pool: node step", find("node step: one ray per lane")), ("pool: leaf step", find("leaf step: FOUR lanes per ray")),
         ("pool: epilogue", find("the slots' new states are visible")), ("cast_kernel wrapper", kernel_line)]
marks = sorted([(nme, l) for nme, l in marks if l], key=lambda t: t[1])
addr = {}   # address -> [IE, TE, samples, best line]
line = None
for r in rows[3:end]:
    if not r:
        continue
    if r[0].strip().isdigit():
        line = int(r[0])
        continue
    a = r[2] if len(r) > 2 else ""
    if not a.startswith("0x"):
        continue
    ie, te, sm = int(r[pos["Instructions Executed"]]), int(r[pos["Thread Instructions Executed"]]), int(r[pos["# Samples"]])
    e = addr.setdefault(a, [ie, te, sm, 0])
    if line is not None and line < kernel_line and line > e[3]:
        e[3] = line
tot = sum(e[0] for e in addr.values()); tsm = sum(e[2] for e in addr.values())
per = {}
for e in addr.values():
    name = "cast_kernel wrapper" if e[3] == 0 else [nme for nme, l in marks if e[3] >= l][-1]
    b = per.setdefault(name, [0, 0, 0]); b[0] += e[0]; b[1] += e[1]; b[2] += e[2]
print(f"unique SASS instructions {len(addr)}  warp-inst {tot/1e6:.1f} M  samples {tsm}")
for nme, l in marks:
    if nme in per:
        b = per[nme]
        print(f"  {nme:34s} (line {l:4d}+) {b[0]/1e6:7.1f} M  {100*b[0]/tot:5.1f}%  samples {100*b[2]/max(tsm,1):5.1f}%  thr/inst {b[1]/max(b[0],1):5.1f}")
