"""ctypes bindings of the two CPU checkers.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package (j3d_b200) never does.

  Oracle   oracle/libj3d_oracle.so   plain-C restatement of the reference algorithm
  Ref      oracle/_ref/libj3d_ref.so the unmodified reference compiled from /root/reference
                                      (prebuilt here; travels to the GPU box as a .so)
"""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from j3d_b200.capi import PICK_DTYPE, PIXEL_DTYPE, View  # noqa: E402  (struct layouts only)

_vp, _u32, _fp = C.c_void_p, C.c_uint32, C.POINTER(C.c_float)


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return _vp(a.ctypes.data)


def ref_available() -> bool:
    return (HERE / "_ref" / "libj3d_ref.so").exists()


class Oracle:
    """Plain-C restatement (j3d_oracle.c)."""

    def __init__(self):
        path = HERE / "libj3d_oracle.so"
        if not path.exists():
            raise RuntimeError(f"{path} missing: run `make -C oracle oracle`")
        L = C.CDLL(str(path))
        L.orc_make_projection.argtypes = [_u32, _u32, _fp, _fp, _fp]
        L.orc_unzoom.argtypes = [_fp, _fp, _fp, _fp, _fp, _fp]
        L.orc_make_matcap.argtypes = [C.c_int, _vp, C.POINTER(_u32)]
        L.orc_fill_background.argtypes = [_u32, _u32, _u32, _u32, _vp]
        L.orc_mesh_create.restype = _vp
        L.orc_mesh_create.argtypes = [_vp, _u32, _vp, _u32, _vp, _vp, _vp, _u32, _u32, _vp, _u32]
        L.orc_mesh_destroy.argtypes = [_vp]
        L.orc_mesh_normals.restype = _vp
        L.orc_mesh_normals.argtypes = [_vp]
        L.orc_mesh_bbox.argtypes = [_vp, _fp, _fp]
        L.orc_find_closest.argtypes = [_vp, _vp, _u32, _vp, _vp]
        L.orc_cast.argtypes = [C.POINTER(_vp), _u32, C.POINTER(View), C.c_int, C.c_int, C.c_int, C.c_int, _vp, _u32]
        L.orc_shade.argtypes = [_vp, _u32, C.POINTER(View), _vp, _u32, _u32, _u32, _vp, _u32]
        L.orc_splat.argtypes = [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_u32), C.POINTER(_vp),
                                C.POINTER(_u32), _u32, C.POINTER(View), _vp, _vp, _u32, _vp, _u32]
        L.orc_find_all.restype = _u32
        L.orc_find_all.argtypes = [_vp, _vp, _u32, _vp, _vp, _vp, _u32]
        L.orc_voxelize.argtypes = [_vp, _u32, C.POINTER(_u32), _vp, _vp]
        L.orc_pick.argtypes = [_vp, _u32, C.POINTER(View), C.POINTER(_vp), _u32, C.POINTER(_vp), C.POINTER(_u32), C.POINTER(_vp),
                               C.POINTER(_u32), _u32, _vp, _u32, _vp]
        self.L = L

    def make_view(self, w, h, bb_min, bb_max, flags) -> View:
        v = View()
        v.width, v.height, v.flags = w, h, flags
        near = C.c_float()
        self.L.orc_make_projection(w, h, C.byref(near), v.projection, v.projection_inv)
        v.near_plane = near.value
        mn = (C.c_float * 3)(*[float(x) for x in bb_min])
        mx = (C.c_float * 3)(*[float(x) for x in bb_max])
        d = C.c_float()
        self.L.orc_unzoom(mn, mx, C.byref(d), v.pivot, v.cs, v.cs_inv)
        v.diagonal = d.value
        return v

    def make_matcap(self, kind=0):
        out = np.empty((512, 512), np.uint32)
        cav = _u32()
        self.L.orc_make_matcap(kind, _p(out), C.byref(cav))
        return out, cav.value

    def fill_background(self, w, h, top=0xFF000000, bottom=0xFF404040):
        out = np.empty((h, w), np.uint32)
        self.L.orc_fill_background(w, h, top, bottom, _p(out))
        return out

    def mesh(self, verts, tris, vcolors=None, uv=None, texture=None, cs=None, db_id=0x20000000):
        tw = th = 0
        if texture is not None:
            th, tw = texture.shape
        csb = None if cs is None else np.ascontiguousarray(cs, np.float32)
        h = self.L.orc_mesh_create(_p(verts), verts.shape[0], _p(tris), tris.shape[0], _p(vcolors), _p(uv), _p(texture), tw, th, _p(csb), db_id)
        return OracleMesh(self, h, tris.shape[0])

    def cast(self, meshes, view: View, rect=None, out=None):
        w, h = view.width, view.height
        if out is None:
            out = np.zeros((h, w), PIXEL_DTYPE)
            out["object_id"] = 0xFFFFFFFF
        x0, y0, x1, y1 = rect if rect is not None else (0, 0, w - 1, h - 1)
        arr = (_vp * max(1, len(meshes)))(*[m.h for m in meshes])
        self.L.orc_cast(arr, len(meshes), C.byref(view), x0, y0, x1, y1, _p(out), w)
        return out

    def shade(self, pixels, view: View, matcap, cavity, rgba):
        """rgba: [H,W] u32 pre-filled with the background; modified in place."""
        mh, mw = matcap.shape
        self.L.orc_shade(_p(pixels), view.width, C.byref(view), _p(matcap), mw, mh, cavity, _p(rgba), view.width)
        return rgba

    def splat(self, clouds, view: View, pixels_in, pixels_inout, rgba):
        """clouds: list of (pos, nrm|None, clr|None, cs|None, db_id)."""
        n = len(clouds)
        pos = (_vp * n)(*[_p(c[0]) for c in clouds])
        nrm = (_vp * n)(*[_p(c[1]) for c in clouds])
        clr = (_vp * n)(*[_p(c[2]) for c in clouds])
        cnt = (_u32 * n)(*[c[0].shape[0] for c in clouds])
        keep = [None if c[3] is None else np.ascontiguousarray(c[3], np.float32) for c in clouds]
        css = (_vp * n)(*[_p(k) for k in keep])
        ids = (_u32 * n)(*[c[4] for c in clouds])
        self.L.orc_splat(pos, nrm, clr, cnt, css, ids, n, C.byref(view), _p(pixels_in), _p(pixels_inout), view.width, _p(rgba), view.width)


    def pick(self, pixels, view: View, meshes, clouds, xy):
        """clouds: list of (pos, cs|None, db_id).  Returns PICK_DTYPE [n]."""
        xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
        out = np.zeros((xy.shape[0],), PICK_DTYPE)
        marr = (_vp * max(1, len(meshes)))(*[m.h for m in meshes])
        n = len(clouds)
        ident = np.eye(4, dtype=np.float32)
        keep = [ident if c[1] is None else np.ascontiguousarray(c[1], np.float32) for c in clouds]
        pos = (_vp * max(1, n))(*[_p(c[0]) for c in clouds])
        cnt = (_u32 * max(1, n))(*[c[0].shape[0] for c in clouds])
        css = (_vp * max(1, n))(*[_p(k) for k in keep])
        ids = (_u32 * max(1, n))(*[c[2] for c in clouds])
        self.L.orc_pick(_p(pixels), view.width, C.byref(view), marr, len(meshes), pos, cnt, css, ids, n, _p(xy), xy.shape[0], _p(out))
        return out


class OracleMesh:
    def __init__(self, orc: Oracle, h, nt):
        self.orc, self.h, self.nt = orc, h, nt

    def normals(self):
        p = self.orc.L.orc_mesh_normals(self.h)
        return np.ctypeslib.as_array(C.cast(p, _fp), shape=(self.nt, 3)).copy()

    def bbox(self):
        mn, mx = (C.c_float * 3)(), (C.c_float * 3)()
        self.orc.L.orc_mesh_bbox(self.h, mn, mx)
        return np.array(mn[:], np.float32), np.array(mx[:], np.float32)

    def find_closest(self, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        hits = np.zeros((n, 4), np.float32)
        ids = np.zeros((n,), np.uint32)
        self.orc.L.orc_find_closest(self.h, _p(rays), n, _p(hits), _p(ids))
        return hits, ids

    def find_all(self, rays):
        """qbvh::find_all_triangles: (offsets [n+1], hits [total,4] {u,v,t,0}, ids [total])."""
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        offsets = np.zeros((n + 1,), np.uint32)
        total = self.orc.L.orc_find_all(self.h, _p(rays), n, _p(offsets), None, None, 0)
        hits = np.zeros((max(total, 1), 4), np.float32)
        ids = np.zeros((max(total, 1),), np.uint32)
        self.orc.L.orc_find_all(self.h, _p(rays), n, _p(offsets), _p(hits), _p(ids), total)
        return offsets, hits[:total], ids[:total]

    def voxelize(self, max_dim):
        """_write_vox's grid: (vmax [Z,Y,X] u8, vmin [Z,Y,X] u8)."""
        dims = (_u32 * 3)()
        self.orc.L.orc_voxelize(self.h, max_dim, dims, None, None)
        shape = (dims[2], dims[1], dims[0])
        vmax, vmin = np.zeros(shape, np.uint8), np.zeros(shape, np.uint8)
        self.orc.L.orc_voxelize(self.h, max_dim, dims, _p(vmax), _p(vmin))
        return vmax, vmin

    def destroy(self):
        if self.h:
            self.orc.L.orc_mesh_destroy(self.h)
            self.h = None


class Ref:
    """The unmodified reference renderer (oracle/_ref/libj3d_ref.so)."""

    def __init__(self, w, h):
        path = HERE / "_ref" / "libj3d_ref.so"
        if not path.exists():
            raise RuntimeError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(str(path))
        L.ref_create.restype = _vp
        L.ref_create.argtypes = [_u32, _u32]
        L.ref_destroy.argtypes = [_vp]
        L.ref_add_mesh.restype = _u32
        L.ref_add_mesh.argtypes = [_vp, _vp, _u32, _vp, _u32, _vp, _vp, _vp, _u32, _u32, _vp]
        L.ref_add_cloud.restype = _u32
        L.ref_add_cloud.argtypes = [_vp, _vp, _vp, _vp, _u32, _vp]
        L.ref_unzoom.argtypes = [_vp]
        L.ref_get_view.argtypes = [_vp, C.POINTER(View)]
        L.ref_set_view.argtypes = [_vp, C.POINTER(View)]
        L.ref_set_matcap.argtypes = [_vp, C.c_int]
        L.ref_get_matcap.argtypes = [_vp, _vp, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)]
        L.ref_render.argtypes = [_vp, _u32]
        L.ref_get_pixels.argtypes = [_vp, C.c_int, _vp]
        L.ref_get_image.argtypes = [_vp, _vp]
        L.ref_get_times.argtypes = [_vp, C.POINTER(C.c_double)]
        L.ref_time_qbvh.restype = C.c_double
        L.ref_time_qbvh.argtypes = [_vp, _u32, _vp, _u32, C.POINTER(_u32)]
        L.ref_find_closest.argtypes = [_vp, _u32, _vp, _u32, _u32, _vp, _u32, _vp, _vp]
        L.ref_hardware_concurrency.restype = C.c_int
        L.ref_pick.argtypes = [_vp, _vp, _u32, _vp]
        L.ref_find_all.restype = _u32
        L.ref_find_all.argtypes = [_vp, _u32, _vp, _u32, _vp, _u32, _vp, _vp, _vp, _u32]
        L.ref_voxelize.argtypes = [_vp, _u32, _vp, _u32, _vp, _vp, _vp, _u32, _u32, _u32, C.POINTER(_u32), _vp, C.c_uint64]
        self.L, self.w, self.h = L, w, h
        self.s = L.ref_create(w, h)

    def close(self):
        if self.s:
            self.L.ref_destroy(self.s)
            self.s = None

    def cores(self):
        return self.L.ref_hardware_concurrency()

    def add_mesh(self, verts, tris, vcolors=None, uv=None, texture=None, cs=None):
        tw = th = 0
        if texture is not None:
            th, tw = texture.shape
        csb = None if cs is None else np.ascontiguousarray(cs, np.float32)
        return self.L.ref_add_mesh(self.s, _p(verts), verts.shape[0], _p(tris), tris.shape[0], _p(vcolors), _p(uv), _p(texture), tw, th, _p(csb))

    def add_cloud(self, pos, nrm=None, clr=None, cs=None):
        csb = None if cs is None else np.ascontiguousarray(cs, np.float32)
        return self.L.ref_add_cloud(self.s, _p(pos), _p(nrm), _p(clr), pos.shape[0], _p(csb))

    def pick(self, xy):
        """Picking through the reference's own code on canvas::_canvas of the last render.  PICK_DTYPE [n]."""
        xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
        out = np.zeros((xy.shape[0],), PICK_DTYPE)
        self.L.ref_pick(self.s, _p(xy), xy.shape[0], _p(out))
        return out

    def find_all(self, verts, tris, rays):
        """The reference's qbvh::find_all_triangles: (offsets [n+1], hits [total,4], ids [total])."""
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        offsets = np.zeros((n + 1,), np.uint32)
        total = self.L.ref_find_all(_p(verts), verts.shape[0], _p(tris), tris.shape[0], _p(rays), n, _p(offsets), None, None, 0)
        hits = np.zeros((max(total, 1), 4), np.float32)
        ids = np.zeros((max(total, 1),), np.uint32)
        self.L.ref_find_all(_p(verts), verts.shape[0], _p(tris), tris.shape[0], _p(rays), n, _p(offsets), _p(hits), _p(ids), total)
        return offsets, hits[:total], ids[:total]

    def voxelize(self, verts, tris, max_dim, vcolors=None, uv=None, texture=None):
        """The reference's write_vox end to end (through a temporary .vox file): grid [Z,Y,X] u8."""
        tw = th = 0
        if texture is not None:
            th, tw = texture.shape
        dims = (_u32 * 3)()
        rc = self.L.ref_voxelize(_p(verts), verts.shape[0], _p(tris), tris.shape[0], _p(vcolors), _p(uv), _p(texture), tw, th, max_dim, dims, None, 0)
        if rc != 0:
            raise RuntimeError(f"ref_voxelize failed: {rc}")
        grid = np.zeros((dims[2], dims[1], dims[0]), np.uint8)
        rc = self.L.ref_voxelize(_p(verts), verts.shape[0], _p(tris), tris.shape[0], _p(vcolors), _p(uv), _p(texture), tw, th, max_dim, dims, _p(grid), grid.size)
        if rc != 0:
            raise RuntimeError(f"ref_voxelize failed: {rc}")
        return grid

    def unzoom(self):
        self.L.ref_unzoom(self.s)

    def view(self) -> View:
        v = View()
        self.L.ref_get_view(self.s, C.byref(v))
        return v

    def set_view(self, v: View):
        self.L.ref_set_view(self.s, C.byref(v))

    def set_matcap(self, kind):
        self.L.ref_set_matcap(self.s, kind)

    def matcap(self):
        w, h, cav = _u32(), _u32(), _u32()
        self.L.ref_get_matcap(self.s, None, C.byref(w), C.byref(h), C.byref(cav))
        out = np.empty((h.value, w.value), np.uint32)
        self.L.ref_get_matcap(self.s, _p(out), C.byref(w), C.byref(h), C.byref(cav))
        return out, cav.value

    def render(self, stages=7):
        self.L.ref_render(self.s, stages)

    def pixels(self, which=0):
        out = np.zeros((self.h, self.w), PIXEL_DTYPE)
        self.L.ref_get_pixels(self.s, which, _p(out))
        return out

    def image(self):
        out = np.zeros((self.h, self.w), np.uint32)
        self.L.ref_get_image(self.s, _p(out))
        return out

    def times(self):
        t = (C.c_double * 5)()
        self.L.ref_get_times(self.s, t)
        return dict(build=t[0], cast=t[1], shade=t[2], splat=t[3], copy=t[4])

    def time_qbvh(self, verts, tris):
        n = _u32()
        t = self.L.ref_time_qbvh(_p(verts), verts.shape[0], _p(tris), tris.shape[0], C.byref(n))
        return t, n.value

    def find_closest(self, verts, tris, rays, leaf_size=32):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        hits = np.zeros((n, 4), np.float32)
        ids = np.zeros((n,), np.uint32)
        self.L.ref_find_closest(_p(verts), verts.shape[0], _p(tris), tris.shape[0], leaf_size, _p(rays), n, _p(hits), _p(ids))
        return hits, ids


# ---- ingest (oracle/ref_ingest.cpp: the unmodified jtk PLY reader and j3d/pc.cpp estimate_normals) -------------------
def _ref_lib():
    path = HERE / "_ref" / "libj3d_ref.so"
    if not path.exists():
        raise RuntimeError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
    L = C.CDLL(str(path))
    L.ref_read_ply.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64), _vp, _vp, _vp, _vp, _vp]
    L.ref_estimate_normals.argtypes = [_vp, _u32, _u32, _vp]
    L.ref_estimate_normals.restype = None
    return L


def ref_read_ply(data: bytes):
    """jtk::read_ply_from_memory -> dict(vertices, normals, colors, triangles, uv), or None when the reader rejects the file."""
    L = _ref_lib()
    counts = (C.c_uint64 * 5)()
    if not L.ref_read_ply(data, len(data), counts, None, None, None, None, None):
        return None
    v = np.zeros((counts[0], 3), np.float32); n = np.zeros((counts[1], 3), np.float32); c = np.zeros((counts[2],), np.uint32)
    t = np.zeros((counts[3], 3), np.uint32); u = np.zeros((counts[4], 6), np.float32)
    L.ref_read_ply(data, len(data), counts, _p(v), _p(n), _p(c), _p(t), _p(u))
    return {"vertices": v, "normals": n, "colors": c, "triangles": t, "uv": u}


def ref_estimate_normals(pos: np.ndarray, k: int) -> np.ndarray:
    """estimate_normals (j3d/pc.cpp:256-334)."""
    pos = np.ascontiguousarray(pos, np.float32)
    out = np.zeros_like(pos)
    _ref_lib().ref_estimate_normals(_p(pos), pos.shape[0], k, _p(out))
    return out
