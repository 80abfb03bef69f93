"""Distribution of per-ray traversal cost (config B): where does the time of the cast kernel go?"""
import sys; sys.path.insert(0, '.')
import numpy as np, j3d_b200 as j
f = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
W, H = 1920, 1080
verts, tris = j.icosphere(f)
ctx = j.Context(0); m = ctx.mesh_create(verts, tris)
mn, mx = j.compute_bb(verts); v0 = j.make_view(W, H, mn, mx)
for ang in (0.0, 48.0):
    n, t = ctx.cast_cost_image([m], j.orbit_view(v0, ang))
    print(f"angle {ang}: nodes mean {n.mean():.2f} max {n.max()}  tris mean {t.mean():.2f} max {t.max()}")
    for q in (50, 90, 99, 99.9, 99.99):
        print(f"  p{q}: nodes {np.percentile(n, q):.0f} tris {np.percentile(t, q):.0f}")
    # per 8x4 tile: the warp's cost is ~ the max over its lanes (and the union of paths)
    tn = n.reshape(H // 4, 4, W // 8, 8).max(axis=(1, 3))
    print(f"  tile-max nodes: mean {tn.mean():.1f} p99 {np.percentile(tn, 99):.0f} max {tn.max()}  sum(tile max)*32/sum = {tn.sum()*32/n.sum():.2f}")
    hist, edges = np.histogram(n, bins=[0, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 1 << 20])
    tot = n.sum()
    for k in range(len(hist)):
        sel = (n >= edges[k]) & (n < edges[k + 1])
        print(f"  nodes in [{edges[k]},{edges[k+1]}): rays {hist[k]:8d}  share of all node visits {n[sel].sum()/tot:.3f}")
