"""k-NN normal estimation probe: python scripts/perf_knn.py [points] [k]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, j3d_b200 as j
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
pos, _, _ = j.cloud(n)
ctx = j.Context(0)
c = ctx.cloud_create(pos)
c.knn_normals(k)
t0 = time.perf_counter(); c.knn_normals(k); print(f"knn_normals n={n} k={k}: {1e3 * (time.perf_counter() - t0):.1f} ms wall (device work + copies of lists and normals to the host)")
