"""Two frames in flight: the same orbit frames rendered by ONE context (frames back to back on one stream) and by TWO
contexts alternating (two streams, the next frame's cast kernel moves in while this frame's longest rays drain).
Usage: [J3DG_CONSUMER_BLOCKS=n] python scripts/perf_overlap.py [f] [nframes]"""
import os, sys, time, zlib
sys.path.insert(0, '.')
import numpy as np, torch, j3d_b200 as j
f = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 200
W, H = 1920, 1080
verts, tris = j.icosphere(f)
NL = int(os.environ.get('LANES', '2'))
ctxs = [j.Context(0) for _ in range(NL)]
m = ctxs[0].mesh_create(verts, tris)
mn, mx = j.compute_bb(verts)
v0 = j.make_view(W, H, mn, mx)
mc, cav = j.make_matcap(0)
for c in ctxs:
    c.set_matcap(mc, cav)
px = [torch.empty((H, W, 32), dtype=torch.uint8, device='cuda') for _ in range(NL)]
rgba = [torch.empty((H, W), dtype=torch.int32, device='cuda') for _ in range(NL)]
views = [j.orbit_view(v0, float(k % 360)) for k in range(nframes + 4)]

def run(two):
    for c in ctxs:
        c.synchronize()
    t0 = time.perf_counter()
    for k, v in enumerate(views):
        if k == 4:
            for c in ctxs:
                c.synchronize()
            t0 = time.perf_counter()
        i = k % NL if two else 0
        ctxs[i].render_frame([m], [], v, pixels_out=px[i], rgba_out=rgba[i])
    for c in ctxs:
        c.synchronize()
    return (time.perf_counter() - t0) * 1e3 / nframes

def crc():
    out = []
    for i in range(NL):
        ctxs[i].render_frame([m], [], views[7], pixels_out=px[i], rgba_out=rgba[i]); ctxs[i].synchronize()
        out.append(zlib.crc32(px[i].cpu().numpy().tobytes()) ^ zlib.crc32(rgba[i].cpu().numpy().tobytes()))
    return out

for rep in range(2):
    one = run(False); two = run(True)
    print(f"lanes={NL} budget={os.environ.get('J3DG_LANE_BUDGET','24')} one ctx {one:.3f} ms/frame ({W*H/one/1e3:.0f} Mrays/s)   {NL} ctx {two:.3f} ms/frame ({W*H/two/1e3:.0f} Mrays/s)   crc {crc()}", flush=True)
