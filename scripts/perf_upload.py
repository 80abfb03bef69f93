"""Host -> device load times: j3dg_mesh_create (config B: 504 MB of pageable vertices + indices, then the BVH build) and
j3dg_cloud_create (pageable positions / normals / colours).  python scripts/perf_upload.py [f] [points]"""
import sys, time, statistics
sys.path.insert(0, '.')
import numpy as np, j3d_b200 as j
f = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000
ctx = j.Context(0)   # the context first, like a host application: it pins its upload ring and preloads kernels in the background
verts, tris = j.icosphere(f)
ts = []
for k in range(4):
    t0 = time.perf_counter(); m = ctx.mesh_create(verts, tris); ctx.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    i = m.info(); t1 = time.perf_counter(); m.destroy(); print(f"  call {k}: wall {ts[-1]:.1f} ms, upload {i.upload_ms:.1f}, build {i.build_ms:.2f}, destroy {1e3 * (time.perf_counter() - t1):.1f}")
print(f"mesh_create f={f} ({tris.shape[0]} triangles, {(verts.nbytes + tris.nbytes) / 1e6:.0f} MB): wall ms {['%.1f' % t for t in ts]}  upload_ms {i.upload_ms:.1f} build_ms {i.build_ms:.2f}")
pos, nrm, clr = j.cloud(npts)
ts = []
for k in range(3):
    t0 = time.perf_counter(); c = ctx.cloud_create(pos, nrm, clr); ctx.synchronize(); ts.append(1e3 * (time.perf_counter() - t0)); c.destroy()
print(f"cloud_create {npts} points ({(pos.nbytes + nrm.nbytes + clr.nbytes) / 1e6:.0f} MB): wall ms {['%.1f' % t for t in ts]}")
