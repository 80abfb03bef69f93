"""Soak test of frames in flight: many short frames on 2-4 contexts, several canvas sizes / meshes (also tiny ones whose
launches are smaller than the machine, so several launches are resident at once), with and without shadows; every
sampled frame is compared with the same frame rendered alone.  python scripts/soak_lanes.py [seconds]"""
import sys, time, zlib
sys.path.insert(0, '.')
import numpy as np, torch, j3d_b200 as j
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(1)
mc, cav = j.make_matcap(0)
ctxs = [j.Context(0) for _ in range(4)]
for c in ctxs:
    c.set_matcap(mc, cav)
t_end = time.time() + budget
total = 0
case = 0
while time.time() < t_end:
    f = int(rng.choice([6, 20, 59, 150, 400]))
    w, h = [(160, 96), (320, 200), (640, 360), (1280, 720), (1920, 1080)][int(rng.integers(0, 5))]
    lanes = int(rng.integers(2, 5))
    flags = j.DEFAULT_FLAGS | (j.SHADOW if rng.random() < 0.4 else 0)
    verts, tris = j.icosphere(f)
    m = ctxs[0].mesh_create(verts, tris)
    mn, mx = j.compute_bb(verts)
    v0 = j.make_view(w, h, mn, mx, flags)
    n = int(rng.integers(50, 400))
    views = [j.orbit_view(v0, float(rng.uniform(0, 360))) for _ in range(n)]
    px = [torch.empty((h, w, 32), dtype=torch.uint8, device='cuda') for _ in range(lanes)]
    keep = {}
    rg = [torch.empty((h, w), dtype=torch.int32, device='cuda') for _ in range(n)]
    for k, v in enumerate(views):
        ctxs[k % lanes].render_frame([m], [], v, pixels_out=px[k % lanes], rgba_out=rg[k])
    for c in ctxs:
        c.synchronize()
    for k in rng.choice(n, size=min(n, 12), replace=False):
        want = torch.empty((h, w), dtype=torch.int32, device='cuda')
        ctxs[0].render_frame([m], [], views[int(k)], pixels_out=px[0], rgba_out=want)
        ctxs[0].synchronize()
        assert torch.equal(want, rg[int(k)]), (case, f, w, h, lanes, int(k))
    assert all(c.status() == 0 for c in ctxs)
    m.destroy()
    total += n
    case += 1
print(f"soak ok: {case} cases, {total} frames on 2-4 contexts, every sampled frame equal to the frame rendered alone")
