"""Per-tile wall time of the cast kernel (counting variant): is the kernel tail-bound?"""
import os, sys; sys.path.insert(0, '.')
os.environ["J3DG_DEBUG_TILE_TIMES"] = "/tmp/tile_times.bin"
import numpy as np, j3d_b200 as j
f = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
W, H = 1920, 1080
verts, tris = j.icosphere(f)
ctx = j.Context(0); m = ctx.mesh_create(verts, tris)
mn, mx = j.compute_bb(verts); v0 = j.make_view(W, H, mn, mx)
for ang in (0.0, 0.0, 48.0):
    n, t = ctx.cast_cost_image([m], j.orbit_view(v0, ang))
    tt = np.fromfile("/tmp/tile_times.bin", dtype=np.uint32).reshape(H, W, 2)
    t0 = tt[::4, ::8, 0].astype(np.int64); t1 = tt[::4, ::8, 1].astype(np.int64)
    base = t0.min(); t0 -= base; t1 -= base
    dur = (t1 - t0) / 1e3
    tn = n.reshape(H // 4, 4, W // 8, 8).max(axis=(1, 3))
    print(f"angle {ang}: kernel span {t1.max()/1e3:.0f} us; last tile START at {t0.max()/1e3:.0f} us; tile dur us mean {dur.mean():.1f} p50 {np.percentile(dur,50):.1f} p99 {np.percentile(dur,99):.1f} max {dur.max():.1f}")
    order = np.argsort(t1.ravel())[::-1][:8]
    for o in order:
        y, x = divmod(o, W // 8)
        print(f"   late tile ({x*8},{y*4}) start {t0.ravel()[o]/1e3:.0f} end {t1.ravel()[o]/1e3:.0f} us dur {dur.ravel()[o]:.0f} us, max nodes/lane {tn.ravel()[o]}, sum nodes {n.reshape(H//4,4,W//8,8)[y,:,x,:].sum()}")
    # how many tiles are still running at time T
    for T in (0.25, 0.5, 0.75, 0.9):
        Tt = T * t1.max()
        print(f"   at {T:.2f} of span: running tiles {(t0 <= Tt).sum() - (t1 <= Tt).sum()}, finished {(t1 <= Tt).mean():.3f}")
    c = np.corrcoef(dur.ravel(), tn.ravel())[0, 1]
    print(f"   corr(tile dur, tile max nodes) = {c:.3f}; us per max-node: {dur.sum()/tn.sum():.2f}")
