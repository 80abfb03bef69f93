"""Per-call wall time of the synchronous j3dg_render_frame with pinned host outputs (debug probe)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch, j3d_b200 as j
W, H = 1920, 1080
f = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
verts, tris = j.icosphere(f)
ctx = j.Context(0)
m = ctx.mesh_create(verts, tris)
mn, mx = j.compute_bb(verts)
v0 = j.make_view(W, H, mn, mx)
mc, cav = j.make_matcap(0); ctx.set_matcap(mc, cav)
hpx = torch.empty((H, W, 32), dtype=torch.uint8).pin_memory()
hrgba = torch.empty((H, W), dtype=torch.int32).pin_memory()
for mode in ("sync", "submit"):
    ts = []
    for k in range(12):
        t0 = time.perf_counter()
        if mode == "sync":
            ctx.render_frame([m], [], j.orbit_view(v0, float(k)), pixels_out=hpx, rgba_out=hrgba)
        else:
            ctx.frame_submit([m], [], j.orbit_view(v0, float(k)), pixels_out=hpx, rgba_out=hrgba); ctx.frame_wait()
        ts.append(1e3 * (time.perf_counter() - t0))
    print(mode, " ".join(f"{t:.2f}" for t in ts))
