"""BVH build timing probe: config B mesh, median of 7 rebuilds from resident data.  [J3DG_LIB=variant.so] python scripts/perf_build.py [f]"""
import os, sys, statistics
sys.path.insert(0, '.')
import j3d_b200 as j
f = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
verts, tris = j.icosphere(f)
ctx = j.Context(0)
m = ctx.mesh_create(verts, tris)
b = []
for k in range(7):
    m.rebuild(); b.append(m.info().build_ms)
print(f"lib={os.environ.get('J3DG_LIB','default')} f={f} build_ms med={statistics.median(b):.3f} min={min(b):.3f} nodes={m.info().nr_of_nodes}", flush=True)
