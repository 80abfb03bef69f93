"""A/B: the cast result must not depend on the tuning (lane budget / algorithm) except at exact ties."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch, j3d_b200 as j
f = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
W, H = 1920, 1080
verts, tris = j.icosphere(f)
ctx = j.Context(0); m = ctx.mesh_create(verts, tris)
mn, mx = j.compute_bb(verts); v0 = j.make_view(W, H, mn, mx)
v = j.orbit_view(v0, 30.0)
def render(budget, algo):
    ctx.set_tuning(budget, algo)
    t = torch.full((H, W, 8), 12345, dtype=torch.int32, device="cuda")  # poison: pixels nobody writes show up
    ctx.cast([m], v, t)
    ctx.synchronize()
    return t.cpu().numpy().view(j.PIXEL_DTYPE).reshape(H, W)
ref = render(1 << 30, 0)
print("ref hits", int((ref["object_id"] != 0xFFFFFFFF).sum()), "poison left", int((ref["object_id"] == 12345).sum() - 0))
for budget, algo in [(1, 0), (4, 0), (16, 0), (32, 0), (64, 0), (0, 1)]:
    a = render(budget, algo)
    d = (a["object_id"] != ref["object_id"])
    hm = ((a["object_id"] == 0xFFFFFFFF) != (ref["object_id"] == 0xFFFFFFFF))
    dd = np.abs(a["depth"][d & ~hm].astype(np.float64) - ref["depth"][d & ~hm]) / np.maximum(1e-30, np.abs(ref["depth"][d & ~hm]))
    print(f"budget={budget} algo={algo}: id diffs {int(d.sum())}, hit/miss diffs {int(hm.sum())}, poison {int((a['object_id']==12345).sum())}, "
          f"max rel depth diff among id-diffs {dd.max() if dd.size else 0:.2e}, depth diffs elsewhere {int((a['depth'][~d] != ref['depth'][~d]).sum())}")
    if hm.sum():
        ys, xs = np.nonzero(hm)
        for y, x in list(zip(ys, xs))[:5]:
            print("    ", x, y, "a:", a[y, x], "ref:", ref[y, x])
