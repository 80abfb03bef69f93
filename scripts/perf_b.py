"""Quick perf probe: config B (28M-triangle icosphere) build + 1080p cast/shade timings."""
import sys, time; sys.path.insert(0,'.')
import numpy as np, j3d_b200 as j
f=int(sys.argv[1]) if len(sys.argv)>1 else 1184
W,H=1920,1080
t=time.time(); verts,tris=j.icosphere(f); print("gen",time.time()-t,tris.shape, flush=True)
ctx=j.Context(0)
t=time.time(); m=ctx.mesh_create(verts,tris); print("create wall",time.time()-t, flush=True)
i=m.info(); print("build_ms",i.build_ms,"upload_ms",i.upload_ms,"nodes",i.nr_of_nodes, "tris/node", i.nr_of_triangles/max(1,i.nr_of_nodes))
for k in range(3):
    m.rebuild(); print("rebuild_ms",m.info().build_ms)
mn,mx=j.compute_bb(verts)
v0=j.make_view(W,H,mn,mx)
mc,cav=j.make_matcap(0); ctx.set_matcap(mc,cav)
import torch
px=torch.empty((H,W,32),dtype=torch.uint8,device='cuda'); rgba=torch.empty((H,W),dtype=torch.int32,device='cuda')
for ang in [0,0,30,60,90,120]:
    v=j.orbit_view(v0,ang)
    ctx.render_frame([m],[],v,pixels_out=px,rgba_out=rgba)
    tm=ctx.timings()
    print("angle",ang,"cast_ms",round(tm.cast_ms,3),"shade_ms",round(tm.shade_ms,3),"Mrays/s",round(W*H/tm.cast_ms/1e3,1))
print("stats nodes/ray, tris/ray",ctx.cast_stats([m],v0))
vs=j.orbit_view(v0,30); vs.flags|=j.SHADOW
ctx.render_frame([m],[],vs,pixels_out=px,rgba_out=rgba); tm=ctx.timings(); print("shadow cast_ms",tm.cast_ms,"rays",tm.rays)
hits=(px.view(torch.int32).reshape(H,W,8)[:,:,4]!=-1).sum().item(); print("hit px",hits)
