"""Configs C and P of BASELINE.json at full size on one GPU (sanity + timings; not bench lines).
  P: 100 M-point vertex-coloured cloud, 1080p depth splat on top of the config-A mesh
  C: 300 M-triangle mesh, 3840x2160, shadow rays
Usage: python scripts/config_cp.py [P] [C]"""
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch, j3d_b200 as j
which = sys.argv[1:] or ["P", "C"]
ctx = j.Context(0)
mc, cav = j.make_matcap(0); ctx.set_matcap(mc, cav)
if "P" in which:
    W, H = 1920, 1080
    t = time.time(); pos, nrm, clr = j.cloud(100_000_000); print(f"P: generated 100M points in {time.time()-t:.1f}s", flush=True)
    verts, tris = j.icosphere(59)
    verts = (verts * 0.8).astype(np.float32)
    m = ctx.mesh_create(verts, tris)
    t = time.time(); cl = ctx.cloud_create(pos, nrm, clr); ctx.synchronize(); print(f"P: cloud upload {time.time()-t:.2f}s", flush=True)
    mn, mx = j.compute_bb(pos)
    v0 = j.make_view(W, H, mn, mx)
    px = torch.empty((H, W, 32), dtype=torch.uint8, device='cuda'); rgba = torch.empty((H, W), dtype=torch.int32, device='cuda')
    for k in range(6):
        ctx.timings(reset=True)
        ctx.render_frame([m], [cl], j.orbit_view(v0, 20.0 * k), pixels_out=px, rgba_out=rgba)
        tm = ctx.timings(reset=True)
        print(f"P: frame {k}: cast {tm.cast_ms:.3f} ms shade {tm.shade_ms:.3f} ms splat {tm.splat_ms:.3f} ms -> {100e6/tm.splat_ms/1e6:.0f} Gpoints/s... ({100e6/(tm.splat_ms*1e-3)/1e9:.1f} Gpts/s), 1.2 GB positions -> {1.2/(tm.splat_ms*1e-3):.0f} GB/s", flush=True)
    p = px.cpu().numpy().view(j.PIXEL_DTYPE).reshape(H, W)
    pts = p["db_id"] == 0x40000000
    print(f"P: pixels covered by points {int(pts.sum())}, by mesh {int((p['db_id']==0x20000000).sum())}, max point id {int(p['object_id'][pts].max())}")
    cl.destroy(); m.destroy(); del pos, nrm, clr
if "C" in which:
    W, H = 3840, 2160
    t = time.time(); verts, tris = j.icosphere(3873); print(f"C: generated {tris.shape[0]} triangles in {time.time()-t:.1f}s", flush=True)
    t = time.time(); m = ctx.mesh_create(verts, tris); ctx.synchronize(); print(f"C: mesh_create wall {time.time()-t:.2f}s", flush=True)
    i = m.info(); print(f"C: build_ms {i.build_ms:.1f} upload_ms {i.upload_ms:.1f} nodes {i.nr_of_nodes}  BVH {(i.nr_of_nodes*i.node_bytes + i.nr_of_triangles*i.triangle_bytes)/1e9:.1f} GB; device mem used {torch.cuda.mem_get_info()[1]/1e9 - torch.cuda.mem_get_info()[0]/1e9:.1f} GB", flush=True)
    mn, mx = j.compute_bb(verts)
    v0 = j.make_view(W, H, mn, mx, j.DEFAULT_FLAGS | j.SHADOW)
    px = torch.empty((H, W, 32), dtype=torch.uint8, device='cuda'); rgba = torch.empty((H, W), dtype=torch.int32, device='cuda')
    for k in range(5):
        ctx.timings(reset=True)
        ctx.render_frame([m], [], j.orbit_view(v0, 25.0 * k), pixels_out=px, rgba_out=rgba)
        tm = ctx.timings(reset=True)
        print(f"C: frame {k}: cast(primary+shadow) {tm.cast_ms:.3f} ms shade {tm.shade_ms:.3f} ms rays {tm.rays} -> {tm.rays/tm.cast_ms/1e3:.0f} Mrays/s", flush=True)
    p = px.cpu().numpy().view(j.PIXEL_DTYPE).reshape(H, W)
    hit = p["object_id"] != 0xFFFFFFFF
    print(f"C: hits {int(hit.sum())} shadowed {int((p['mark'][hit] & 1).sum())} max tri id {int(p['object_id'][hit].max())}")
    bu, bv = p["barycentric_u"][hit], p["barycentric_v"][hit]
    print("C: barycentrics ok", bool((bu >= -1e-5).all() and (bv >= -1e-5).all() and (bu + bv <= 1 + 1e-5).all()))
