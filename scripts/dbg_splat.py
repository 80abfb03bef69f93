import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, j3d_b200 as j
from oracle.bindings import Oracle
orc=Oracle(); ctx=j.Context(0)
w,h=320,180
verts,tris=j.icosphere(8); verts=(verts*0.6).astype(np.float32)
for n,flags in [(200003,j.DEFAULT_FLAGS),(50001,j.DEFAULT_FLAGS|j.ONE_BIT)]:
    pos,nrm,clr=j.cloud(n)
    mn,mx=j.compute_bb(np.concatenate([verts,pos]))
    v=j.orbit_view(j.make_view(w,h,mn,mx,flags),15.0)
    om=orc.mesh(verts,tris); px=orc.cast([om],v)
    mc,cav=j.make_matcap(0)
    rgba0=orc.shade(px,v,mc,cav,j.fill_background(w,h))
    wp,wr=px.copy(),rgba0.copy()
    orc.splat([(pos,nrm,clr,None,0x40000000)],v,px,wp,wr)
    cl=ctx.cloud_create(pos,nrm,clr)
    gp,gr=px.copy(),rgba0.copy()
    ctx.splat([cl],v,px,gp,gr)
    idm=(gp['object_id']!=wp['object_id'])|(gp['db_id']!=wp['db_id'])
    dm=(gp['depth']!=wp['depth'])&~idm
    cm=(gr!=wr)&~idm
    print(n,hex(flags),"pts px",int((wp['db_id']==0x40000000).sum()),"idm",int(idm.sum()),"depth mm",int(dm.sum()),"rgba mm",int(cm.sum()))
    ys,xs=np.nonzero(dm)
    for y,x in list(zip(ys,xs))[:5]:
        print("  d",y,x,gp[y,x],wp[y,x], px[y,x])
    ys,xs=np.nonzero(idm)
    for y,x in list(zip(ys,xs))[:5]:
        print("  id",y,x,gp[y,x],wp[y,x],hex(gr[y,x]),hex(wr[y,x]))
