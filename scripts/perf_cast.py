"""Cast-kernel tuning probe: config B (28M-triangle icosphere), 1080p, orbit frames.
Usage: [J3DG_LIB=path/to/variant.so] python scripts/perf_cast.py [f] [nframes]
Prints median / min cast_ms over the orbit frames, the BVH stats and a checksum of the pixel
buffer (so variants can be compared for identical output)."""
import os, sys, time, statistics, zlib
sys.path.insert(0, '.')
import numpy as np, torch, j3d_b200 as j
f = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 24
W, H = 1920, 1080
verts, tris = j.icosphere(f)
ctx = j.Context(0)
m = ctx.mesh_create(verts, tris)
i = m.info()
b = []
for k in range(3):
    m.rebuild(); b.append(m.info().build_ms)
mn, mx = j.compute_bb(verts)
v0 = j.make_view(W, H, mn, mx)
mc, cav = j.make_matcap(0); ctx.set_matcap(mc, cav)
px = torch.empty((H, W, 32), dtype=torch.uint8, device='cuda'); rgba = torch.empty((H, W), dtype=torch.int32, device='cuda')
for ang in range(3):
    ctx.render_frame([m], [], j.orbit_view(v0, float(ang)), pixels_out=px, rgba_out=rgba)
ctx.timings(reset=True)
cast = []
for k in range(nframes):
    ctx.render_frame([m], [], j.orbit_view(v0, float(3 + k * 15 % 360)), pixels_out=px, rgba_out=rgba)
    tm = ctx.timings(reset=True)
    cast.append(tm.cast_ms)
ctx.render_frame([m], [], j.orbit_view(v0, 30.0), pixels_out=px, rgba_out=rgba)
ctx.synchronize()  # the library runs on its own stream
p = px.cpu().numpy().view(j.PIXEL_DTYPE).reshape(H, W)
hit = p["object_id"] != 0xFFFFFFFF
chk = zlib.crc32(p["object_id"].tobytes()) ^ zlib.crc32(p["depth"].tobytes())
vs = j.orbit_view(v0, 30.0); vs.flags |= j.SHADOW
ctx.render_frame([m], [], vs, pixels_out=px, rgba_out=rgba); sh = ctx.timings(reset=True).cast_ms
npr, tpr = ctx.cast_stats([m], v0)
med = statistics.median(cast)
print(f"lib={os.environ.get('J3DG_LIB','default')} f={f} build_ms={statistics.median(b):.2f} nodes={i.nr_of_nodes} "
      f"cast_ms med={med:.3f} min={min(cast):.3f} max={max(cast):.3f} Mrays/s={W*H/med/1e3:.0f} shadow_cast_ms={sh:.3f} "
      f"nodes/ray={npr:.2f} tris/ray={tpr:.2f} hits={int(hit.sum())} crc={chk:08x}", flush=True)
