"""Voxel export timing probe: config B mesh, j3dg_mesh_voxelize into a device grid.  [J3DG_LIB=variant.so] python scripts/perf_vox.py [max_dim]"""
import os, sys, time, statistics
sys.path.insert(0, '.')
import torch, j3d_b200 as j
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 512
verts, tris = j.icosphere(1184)
ctx = j.Context(0)
m = ctx.mesh_create(verts, tris)
grid = torch.empty((dim ** 3 + 64,), dtype=torch.uint8, device="cuda")
ts = []
for k in range(6):
    t0 = time.perf_counter(); m.voxelize(dim, out=grid); ts.append(1e3 * (time.perf_counter() - t0))
print(f"lib={os.environ.get('J3DG_LIB','default')} max_dim={dim} voxelize_ms med={statistics.median(ts[1:]):.3f} min={min(ts):.3f} occupied={int((grid[:dim**3] != 0).sum())}", flush=True)
