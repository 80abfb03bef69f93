import sys; sys.path.insert(0,'.')
import numpy as np, torch, j3d_b200 as j
ctx=j.Context(0); mc,cav=j.make_matcap(0); ctx.set_matcap(mc,cav)
pos,nrm,clr=j.cloud(100_000_000); verts,tris=j.icosphere(59); verts=(verts*0.8).astype(np.float32)
m=ctx.mesh_create(verts,tris); cl=ctx.cloud_create(pos,nrm,clr)
mn,mx=j.compute_bb(pos); v0=j.make_view(1920,1080,mn,mx)
px=torch.empty((1080,1920,32),dtype=torch.uint8,device='cuda'); rgba=torch.empty((1080,1920),dtype=torch.int32,device='cuda')
for k in range(3):
    ctx.render_frame([m],[cl],j.orbit_view(v0,20.0*k),pixels_out=px,rgba_out=rgba)
ctx.synchronize()
