"""CPU restatement of the reference's INGEST path in numpy.  TEST INFRASTRUCTURE, NOT PRODUCT CODE:
only tests/ may import it.  Pinned on the unmodified reference (oracle/_ref: jtk/ply.h + j3d/pc.cpp compiled as they
are) by tests/test_ingest.py, live where libj3d_ref.so exists and through tests/golden/ingest.npz everywhere.

  read_ply_binary    jtk::read_ply (jtk/ply.h:577-690) for the binary storage modes
  knn                jtk::point_tree::find_k_nearest (jtk/point_tree.h:436-475) by brute force
  fit_normals        jtk::fit_plane (jtk/fitting.h:197-215) on those neighbours
  orient             the propagation of estimate_normals (j3d/pc.cpp:284-333), jtk::hashed_heap's queue discipline
"""
from __future__ import annotations

import numpy as np

_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
          "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def read_ply_binary(data: bytes):
    """-> dict(vertices [nv,3] f32, normals, colors [nv] u32 0xAABBGGRR, triangles [nf,3] u32, uv [nf,6] f32); empty
    arrays for what the file lacks (ply.h:598-679).  Values go through a double like rply's ply_get_argument_value."""
    end = data.index(b"end_header")
    end = data.index(b"\n", end) + 1
    lines = data[:end].decode("ascii").splitlines()
    assert lines[0].strip() == "ply"
    order = "<"
    elems = []
    for ln in lines[1:]:
        w = ln.split()
        if not w or w[0] in ("comment", "obj_info", "end_header"):
            continue
        if w[0] == "format":
            assert w[1].startswith("binary")
            order = "<" if w[1] == "binary_little_endian" else ">"
        elif w[0] == "element":
            elems.append((w[1], int(w[2]), []))
        elif w[0] == "property":
            if w[1] == "list":
                elems[-1][2].append((w[4], _TYPES[w[3]], _TYPES[w[2]]))
            else:
                elems[-1][2].append((w[2], _TYPES[w[1]], None))
    out = {"vertices": np.zeros((0, 3), np.float32), "normals": np.zeros((0, 3), np.float32), "colors": np.zeros((0,), np.uint32),
           "triangles": np.zeros((0, 3), np.uint32), "uv": np.zeros((0, 6), np.float32)}
    pos = end
    for name, count, props in elems:
        if name == "vertex":
            dt = np.dtype([(p[0], order + p[1]) for p in props])
            rec = np.frombuffer(data, dt, count, pos)
            pos += dt.itemsize * count
            names = rec.dtype.names
            if all(a in names for a in "xyz"):
                out["vertices"] = np.stack([rec[a].astype(np.float64).astype(np.float32) for a in "xyz"], 1)
            if all(a in names for a in ("nx", "ny", "nz")):
                out["normals"] = np.stack([rec[a].astype(np.float64).astype(np.float32) for a in ("nx", "ny", "nz")], 1)
            chan = []
            for alts in (("red", "r", "diffuse_red"), ("green", "g", "diffuse_green"), ("blue", "b", "diffuse_blue"), ("alpha", "a", "diffuse_alpha")):
                chan.append(next((a for a in alts if a in names), None))
            if chan[0] is not None:  # ply.h:643-644: sized by the red channel
                c = np.full(count, 0xFFFFFFFF, np.uint32)
                for j, a in enumerate(chan):
                    if a is not None:
                        v = rec[a].astype(np.float64).astype(np.int64).astype(np.uint32) & 0xFF
                        c = (c & ~np.uint32(0xFF << (8 * j))) | (v << np.uint32(8 * j))
                out["colors"] = c.astype(np.uint32)
        elif name == "face":
            tris = np.zeros((count, 3), np.uint32)
            uv = np.zeros((count, 6), np.float32)
            have_idx = have_uv = False
            for f in range(count):
                idx_done = uv_done = False
                for pname, ty, cty in props:
                    if cty is None:
                        pos += np.dtype(ty).itemsize
                        continue
                    n = int(np.frombuffer(data, order + cty, 1, pos)[0])
                    pos += np.dtype(cty).itemsize
                    vals = np.frombuffer(data, order + ty, n, pos).astype(np.float64)
                    pos += np.dtype(ty).itemsize * n
                    if pname in ("vertex_indices", "vertex_index") and not idx_done:
                        tris[f] = vals[:3].astype(np.int64).astype(np.uint32)
                        idx_done = have_idx = True
                    elif pname == "texcoord" and not uv_done:
                        m = min(n, 6)
                        uv[f, :m] = vals[:m].astype(np.float32)
                        uv_done = have_uv = True
            if have_idx:
                out["triangles"] = tris
            if have_uv:
                out["uv"] = uv
            break
        else:
            dt = np.dtype([(p[0], order + p[1]) for p in props])
            pos += dt.itemsize * count
    return out


def knn(pos: np.ndarray, k: int) -> np.ndarray:
    """[n, min(k, n)] indices by ascending (distance, index); distances rounded like point_tree.h:52-57."""
    pos = np.ascontiguousarray(pos, np.float32)
    n = pos.shape[0]
    kk = min(k, n)
    out = np.zeros((n, kk), np.uint32)
    for i in range(n):
        d = pos[i] - pos
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]  # float32 throughout
        out[i] = np.lexsort((np.arange(n), d2))[:kk]
    return out


def fit_normals(pos: np.ndarray, nb: np.ndarray):
    """fit_plane per row of nb: float32 centroid / scatter in list order, eigenvector of the eigenvalue of smallest
    magnitude (double eigh of the float matrix).  -> (normals [n,3] f32 (sign arbitrary), eigenvalue gap ratio [n])."""
    pos = np.ascontiguousarray(pos, np.float32)
    n, k = nb.shape
    o = pos[nb[:, 0]].copy()
    for t in range(1, k):
        o = o + pos[nb[:, t]]
    o = o / np.float32(k)
    s = np.zeros((n, 3, 3), np.float32)
    for t in range(k):
        d = pos[nb[:, t]] - o
        s = s + d[:, :, None] * d[:, None, :]
    w, v = np.linalg.eigh(s.astype(np.float64))
    e = np.argmin(np.abs(w), axis=1)
    nrm = v[np.arange(n), :, e]
    ws = np.sort(np.abs(w), axis=1)
    gap = (ws[:, 1] - ws[:, 0]) / np.maximum(ws[:, 2], 1e-300)
    return nrm.astype(np.float32), gap


def orient(nrm: np.ndarray, nb: np.ndarray) -> np.ndarray:
    """pc.cpp:284-333 with the binary min-heap of containers.h:21-92 (strict comparisons, right child on ties)."""
    nrm = nrm.copy()
    n = nrm.shape[0]
    treated = np.zeros(n, bool)
    heap = []  # (score, v0, v1)

    def dot(a, b):
        x = nrm[a] * nrm[b]
        return np.float32(np.float32(x[0] + x[1]) + x[2])

    def push(item):
        heap.append(item)
        i = len(heap) - 1
        while i > 0:
            p = (i - 1) // 2
            if not (item[0] < heap[p][0]):
                break
            heap[i] = heap[p]
            i = p
        heap[i] = item

    def pop():
        top = heap[0]
        last = heap.pop()
        if heap:
            i, ln = 0, len(heap)
            while True:
                l, r = 2 * i + 1, 2 * i + 2
                if r < ln:
                    c = l if heap[l][0] < heap[r][0] else r
                elif l < ln:
                    c = l
                else:
                    break
                if not (heap[c][0] < last[0]):
                    break
                heap[i] = heap[c]
                i = c
            heap[i] = last
        return top

    def push_neighbours(v):
        for u in nb[v]:
            u = int(u)
            if u != v and not treated[u]:
                push((abs(dot(v, u)), v, u))

    v = 0
    while True:
        while v < n and treated[v]:
            v += 1
        if v == n:
            break
        treated[v] = True
        push_neighbours(v)
        while heap:
            _, v0, v1 = pop()
            if treated[v1]:
                continue
            treated[v1] = True
            if dot(v0, v1) < 0:
                nrm[v1] = -nrm[v1]
            push_neighbours(v1)
    return nrm
