"""ctypes binding of the C ABI (include/j3dg.h) plus the host-math helpers of
libj3dg_host.so and the procedural input generators of libj3d_synth.so.

This is the Python face of the *product*: it loads libj3dg.so (hand-written sm_100a CUDA)
and fails loudly if that library is missing or no B200-class GPU is present.  Nothing in
here touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent

ONE_BIT, SHADOW, EDGES, WIREFRAME, SHADING, TEXTURED, VERTEXCOLORS = (1 << i for i in range(7))
STATUS_WAIT_TIMEOUT, STATUS_STACK_OVERFLOW = 1, 2
DEFAULT_FLAGS = EDGES | SHADING | TEXTURED | VERTEXCOLORS

PIXEL_DTYPE = np.dtype(
    [("mark", "u1"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("u", "<f4"), ("v", "<f4"), ("depth", "<f4"),
     ("object_id", "<u4"), ("barycentric_u", "<f4"), ("barycentric_v", "<f4"), ("db_id", "<u4")]
)
assert PIXEL_DTYPE.itemsize == 32
# j3dg_pick (include/j3dg.h)
PICK_DTYPE = np.dtype([("pixel", PIXEL_DTYPE), ("world_pos", "<f4", (3,)), ("closest_vertex", "<u4"), ("pivot", "<f4", (3,)), ("db_id", "<u4")])
assert PICK_DTYPE.itemsize == 64


class View(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("near_plane", C.c_float), ("diagonal", C.c_float),
        ("projection", C.c_float * 16), ("projection_inv", C.c_float * 16), ("cs", C.c_float * 16),
        ("cs_inv", C.c_float * 16), ("pivot", C.c_float * 3), ("flags", C.c_uint32),
    ]

    def copy(self) -> "View":
        v = View()
        C.memmove(C.byref(v), C.byref(self), C.sizeof(View))
        return v


class MeshInfo(C.Structure):
    _fields_ = [
        ("nr_of_vertices", C.c_uint32), ("nr_of_triangles", C.c_uint32), ("nr_of_nodes", C.c_uint32),
        ("nr_of_leaf_triangles", C.c_uint32), ("node_bytes", C.c_uint32), ("triangle_bytes", C.c_uint32),
        ("build_ms", C.c_float), ("upload_ms", C.c_float), ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3),
        ("sah_cost", C.c_float),
    ]


class PlyInfo(C.Structure):
    _fields_ = [("nr_of_vertices", C.c_uint32), ("nr_of_faces", C.c_uint32), ("has_normals", C.c_uint32), ("has_colors", C.c_uint32),
                ("has_uv", C.c_uint32), ("format", C.c_uint32), ("header_bytes", C.c_uint64), ("file_bytes", C.c_uint64),
                ("upload_ms", C.c_float), ("decode_ms", C.c_float)]


class Timings(C.Structure):
    _fields_ = [("cast_ms", C.c_float), ("shade_ms", C.c_float), ("splat_ms", C.c_float), ("copy_ms", C.c_float),
                ("cast_count", C.c_uint32), ("shade_count", C.c_uint32), ("splat_count", C.c_uint32),
                ("kernel_launches", C.c_uint32), ("rays", C.c_uint64)]


_vp = C.c_void_p
_u32 = C.c_uint32
_fp = C.POINTER(C.c_float)


def _ptr(a):
    """numpy array / int device pointer / None -> c_void_p"""
    if a is None:
        return None
    if isinstance(a, int):
        return _vp(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return _vp(a.ctypes.data)
    if hasattr(a, "data_ptr"):  # torch tensor (device or pinned host)
        return _vp(a.data_ptr())
    raise TypeError(type(a))


_lib = None
_host = None
_synth = None


def lib() -> C.CDLL:
    """The CUDA product library.  No fallback: raises if it has not been built."""
    global _lib
    if _lib is None:
        # J3DG_LIB: developer knob to load another BUILD of the same CUDA library (kernel tuning variants)
        path = Path(os.environ["J3DG_LIB"]) if os.environ.get("J3DG_LIB") else PKG / "libj3dg.so"
        if not path.exists():
            raise RuntimeError(f"{path} is missing: run `python -m j3d_b200.build` (nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(str(path), mode=C.RTLD_GLOBAL)
        L.j3dg_last_error.restype = C.c_char_p
        L.j3dg_last_error.argtypes = [_vp]
        L.j3dg_ctx_create.argtypes = [C.c_int, C.POINTER(_vp)]
        L.j3dg_ctx_destroy.argtypes = [_vp]
        L.j3dg_ctx_destroy.restype = None
        L.j3dg_ctx_set_stream.argtypes = [_vp, _vp]
        L.j3dg_ctx_synchronize.argtypes = [_vp]
        L.j3dg_ctx_timings.argtypes = [_vp, C.POINTER(Timings), C.c_int]
        L.j3dg_ctx_set_profiling.argtypes = [_vp, C.c_int]
        L.j3dg_ctx_status.argtypes = [_vp, C.POINTER(_u32), C.c_int]
        L.j3dg_mesh_create.argtypes = [_vp, _vp, _u32, _vp, _u32, _vp, _vp, _vp, _u32, _u32, _u32, _vp, _u32, C.POINTER(_vp)]
        L.j3dg_mesh_destroy.argtypes = [_vp]
        L.j3dg_mesh_destroy.restype = None
        L.j3dg_mesh_rebuild.argtypes = [_vp]
        L.j3dg_mesh_info_get.argtypes = [_vp, C.POINTER(MeshInfo)]
        L.j3dg_mesh_set_cs.argtypes = [_vp, _vp]
        L.j3dg_mesh_bvh_buffer.argtypes = [_vp, C.c_int, C.POINTER(_vp), C.POINTER(C.c_size_t)]
        L.j3dg_mesh_create_empty.argtypes = [_vp, _u32, _u32, _u32, _vp, _u32, C.POINTER(_vp)]
        L.j3dg_mesh_find_closest.argtypes = [_vp, _vp, _u32, _vp, _vp]
        L.j3dg_cast.argtypes = [_vp, C.POINTER(_vp), _u32, C.POINTER(View), C.c_int, C.c_int, C.c_int, C.c_int, _vp, _u32]
        L.j3dg_shade.argtypes = [_vp, _vp, _u32, C.POINTER(View), _vp, _u32, _u32, _u32, _u32, _vp, _vp, _u32]
        L.j3dg_cloud_create.argtypes = [_vp, _vp, _vp, _vp, _u32, _vp, _u32, C.POINTER(_vp)]
        L.j3dg_cloud_destroy.argtypes = [_vp]
        L.j3dg_cloud_destroy.restype = None
        L.j3dg_splat.argtypes = [_vp, C.POINTER(_vp), _u32, C.POINTER(View), _vp, _vp, _u32, _vp, _u32]
        L.j3dg_render_frame.argtypes = [_vp, C.POINTER(_vp), _u32, C.POINTER(_vp), _u32, C.POINTER(View),
                                        _vp, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp]
        L.j3dg_frame_submit.argtypes = L.j3dg_render_frame.argtypes
        L.j3dg_frame_wait.argtypes = [_vp]
        L.j3dg_ctx_set_matcap.argtypes = [_vp, _vp, _u32, _u32, _u32, _u32]
        L.j3dg_ctx_set_tuning.argtypes = [_vp, _u32, C.c_int]
        L.j3dg_ctx_set_screen_shard.argtypes = [_vp, _u32, _u32]
        L.j3dg_cast_stats.argtypes = [_vp, C.POINTER(_vp), _u32, C.POINTER(View), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.j3dg_cast_cost_image.argtypes = [_vp, C.POINTER(_vp), _u32, C.POINTER(View), _vp, _vp]
        L.j3dg_mesh_find_all.argtypes = [_vp, _vp, _u32, _vp, _vp, _vp, _u32, C.POINTER(_u32)]
        L.j3dg_mesh_voxel_dims.argtypes = [_vp, _u32, C.POINTER(_u32)]
        L.j3dg_mesh_voxelize.argtypes = [_vp, _u32, C.POINTER(_u32), _vp, C.c_size_t]
        L.j3dg_ctx_set_dirty_rect.argtypes = [_vp, C.c_int]
        L.j3dg_ctx_readback_bytes.argtypes = [_vp, C.POINTER(C.c_uint64), C.c_int]
        L.j3dg_peer_alloc.argtypes = [_vp, C.c_size_t, C.POINTER(_vp), _vp]
        L.j3dg_peer_free.argtypes = [_vp, _vp]
        L.j3dg_peer_open.argtypes = [_vp, _vp, C.POINTER(_vp)]
        L.j3dg_peer_close.argtypes = [_vp, _vp]
        L.j3dg_stream_signal.argtypes = [_vp, _vp, _u32]
        L.j3dg_stream_wait_geq.argtypes = [_vp, _vp, _u32, _u32]
        L.j3dg_stream_wait_status.argtypes = [_vp, C.POINTER(C.c_int)]
        L.j3dg_pick.argtypes = [_vp, C.POINTER(_vp), _u32, C.POINTER(_vp), _u32, C.POINTER(View), _vp, _u32, _vp, _u32, _vp]
        L.j3dg_group_unique_id.argtypes = [_vp]
        L.j3dg_group_create.argtypes = [_vp, C.c_int, C.c_int, _vp, C.POINTER(_vp)]
        L.j3dg_group_destroy.argtypes = [_vp]
        L.j3dg_group_destroy.restype = None
        L.j3dg_group_rank.argtypes = [_vp]
        L.j3dg_group_world.argtypes = [_vp]
        L.j3dg_group_barrier.argtypes = [_vp]
        L.j3dg_group_max_float.argtypes = [_vp, _vp, _u32]
        L.j3dg_group_allreduce_max_u64.argtypes = [_vp, _vp, C.c_size_t]
        L.j3dg_group_broadcast_mesh.argtypes = [_vp, C.c_int, C.POINTER(_vp)]
        L.j3dg_frames_create.argtypes = [_vp, _u32, _u32, C.c_int, C.c_int, C.POINTER(_vp)]
        L.j3dg_frames_destroy.argtypes = [_vp]
        L.j3dg_frames_destroy.restype = None
        L.j3dg_frames_create_n.argtypes = [_vp, _u32, _u32, C.c_int, C.c_int, _u32, C.POINTER(_vp)]
        L.j3dg_frames_set_lane.argtypes = [_vp, C.c_int, _vp]
        L.j3dg_frames_begin.argtypes = [_vp, C.POINTER(_u32)]
        L.j3dg_frames_target.argtypes = [_vp, _u32, C.POINTER(_vp)]
        L.j3dg_frames_arrive.argtypes = [_vp, _u32]
        L.j3dg_frames_release.argtypes = [_vp, _u32]
        L.j3dg_frames_view.argtypes = [_vp, _u32, C.POINTER(_vp)]
        L.j3dg_ply_decode.argtypes = [_vp, _vp, C.c_size_t, C.POINTER(_vp)]
        L.j3dg_ply_destroy.argtypes = [_vp]
        L.j3dg_ply_destroy.restype = None
        L.j3dg_ply_info_get.argtypes = [_vp, C.POINTER(PlyInfo)]
        L.j3dg_ply_arrays.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]
        L.j3dg_ply_copy.argtypes = [_vp, C.c_int, _vp, C.c_size_t]
        L.j3dg_mesh_create_from_ply.argtypes = [_vp, _vp, _vp, _u32, C.POINTER(_vp)]
        L.j3dg_cloud_create_from_ply.argtypes = [_vp, _vp, _vp, _u32, C.POINTER(_vp)]
        L.j3dg_cloud_estimate_normals.argtypes = [_vp, _u32, _vp]
        L.j3dg_cloud_knn_normals.argtypes = [_vp, _u32, _vp, _vp]
        _lib = L
    return _lib


def host() -> C.CDLL:
    global _host
    if _host is None:
        path = PKG / "libj3dg_host.so"
        if not path.exists():
            raise RuntimeError(f"{path} is missing: run `python -m j3d_b200.build`")
        H = C.CDLL(str(path))
        H.j3dgh_make_projection.argtypes = [_u32, _u32, _fp, _fp, _fp]
        H.j3dgh_invert_orthonormal.argtypes = [_fp, _fp]
        H.j3dgh_matrix_multiply.argtypes = [_fp, _fp, _fp]
        H.j3dgh_unzoom.argtypes = [_fp, _fp, _fp, _fp, _fp, _fp]
        H.j3dgh_orbit.argtypes = [_fp, _fp, C.c_float, _fp, _fp]
        H.j3dgh_make_matcap.argtypes = [C.c_int, _vp, C.POINTER(_u32)]
        H.j3dgh_fill_background.argtypes = [_u32, _u32, _u32, _u32, _u32, _vp]
        H.j3dgh_compute_bb.argtypes = [_vp, _u32, _fp, _fp]
        H.j3dgh_transform_bbox.argtypes = [_fp, _fp, _fp, _fp, _fp]
        _host = H
    return _host


def synth() -> C.CDLL:
    global _synth
    if _synth is None:
        path = PKG / "libj3d_synth.so"
        if not path.exists():
            raise RuntimeError(f"{path} is missing: run `python -m j3d_b200.build`")
        S = C.CDLL(str(path))
        S.synth_icosphere_counts.argtypes = [_u32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        S.synth_icosphere.argtypes = [_u32, C.c_float, _u32, _vp, _vp]
        S.synth_shuffle_triangles.argtypes = [_vp, C.c_uint64, C.c_uint64]
        S.synth_cloud_range.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_float, _vp, _vp, _vp]
        S.synth_vertex_colors.argtypes = [_vp, C.c_uint64, _u32, _vp]
        _synth = S
    return _synth


# ---------------------------------------------------------------------------------------
# procedural inputs
# ---------------------------------------------------------------------------------------
def icosphere(f: int, noise: float = 0.05, seed: int = 1234, shuffle_seed: int | None = None):
    """Noised geodesic icosphere: (vertices [V,3] f32, triangles [T,3] u32), T = 20 f^2."""
    S = synth()
    nv, nt = C.c_uint64(), C.c_uint64()
    S.synth_icosphere_counts(f, C.byref(nv), C.byref(nt))
    verts = np.empty((nv.value, 3), np.float32)
    tris = np.empty((nt.value, 3), np.uint32)
    S.synth_icosphere(f, noise, seed, _ptr(verts), _ptr(tris))
    if shuffle_seed is not None:
        S.synth_shuffle_triangles(_ptr(tris), nt.value, shuffle_seed)
    return verts, tris


def cloud(n: int, seed: int = 42, noise: float = 0.05, first: int = 0):
    S = synth()
    pos = np.empty((n, 3), np.float32)
    nrm = np.empty((n, 3), np.float32)
    clr = np.empty((n,), np.uint32)
    S.synth_cloud_range(first, n, seed, noise, _ptr(pos), _ptr(nrm), _ptr(clr))
    return pos, nrm, clr


def vertex_colors(verts: np.ndarray, seed: int = 7) -> np.ndarray:
    out = np.empty_like(verts)
    synth().synth_vertex_colors(_ptr(verts), verts.shape[0], seed, _ptr(out))
    return out


# ---------------------------------------------------------------------------------------
# host math (camera / pose / matcap), one implementation shared with the C++ canvas mirror
# ---------------------------------------------------------------------------------------
def _f16():
    return (C.c_float * 16)()


def make_view(width: int, height: int, bb_min, bb_max, flags: int = DEFAULT_FLAGS) -> View:
    """Default camera (camera.cpp:5-16) + unzoom pose (scene.cpp:91-111) for one bbox."""
    H = host()
    v = View()
    v.width, v.height, v.flags = width, height, flags
    near = C.c_float()
    H.j3dgh_make_projection(width, height, C.byref(near), v.projection, v.projection_inv)
    v.near_plane = near.value
    mn = (C.c_float * 3)(*[float(x) for x in bb_min])
    mx = (C.c_float * 3)(*[float(x) for x in bb_max])
    diag = C.c_float()
    H.j3dgh_unzoom(mn, mx, C.byref(diag), v.pivot, v.cs, v.cs_inv)
    v.diagonal = diag.value
    return v


def orbit_view(v0: View, angle_deg: float) -> View:
    v = v0.copy()
    host().j3dgh_orbit(v0.cs_inv, v0.pivot, float(angle_deg), v.cs, v.cs_inv)
    return v


def make_matcap(kind: int = 0):
    out = np.empty((512, 512), np.uint32)
    cav = C.c_uint32()
    host().j3dgh_make_matcap(kind, _ptr(out), C.byref(cav))
    return out, cav.value


def fill_background(w: int, h: int, top: int = 0xFF000000, bottom: int = 0xFF404040) -> np.ndarray:
    out = np.empty((h, w), np.uint32)
    host().j3dgh_fill_background(w, h, w, top, bottom, _ptr(out))
    return out


def compute_bb(verts: np.ndarray):
    mn, mx = (C.c_float * 3)(), (C.c_float * 3)()
    host().j3dgh_compute_bb(_ptr(verts), verts.shape[0], mn, mx)
    return np.array(mn[:], np.float32), np.array(mx[:], np.float32)


# ---------------------------------------------------------------------------------------
# object wrappers over the C ABI
# ---------------------------------------------------------------------------------------
class J3dgError(RuntimeError):
    pass


class Context:
    def __init__(self, device: int = 0):
        self._L = lib()
        self._h = _vp()
        rc = self._L.j3dg_ctx_create(device, C.byref(self._h))
        if rc != 0:
            raise J3dgError(f"j3dg_ctx_create({device}) = {rc}: {self._L.j3dg_last_error(None).decode()}")

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise J3dgError(f"{what} = {rc}: {self._L.j3dg_last_error(self._h).decode()}")

    def close(self):
        if self._h:
            self._L.j3dg_ctx_destroy(self._h)
            self._h = _vp()

    def set_stream(self, stream_ptr: int):
        self._check(self._L.j3dg_ctx_set_stream(self._h, _vp(stream_ptr)), "j3dg_ctx_set_stream")

    def synchronize(self):
        self._check(self._L.j3dg_ctx_synchronize(self._h), "j3dg_ctx_synchronize")

    def status(self, reset: bool = False) -> int:
        """Sticky status bits (STATUS_WAIT_TIMEOUT | STATUS_STACK_OVERFLOW), polled from mapped memory without a sync."""
        f = _u32()
        self._check(self._L.j3dg_ctx_status(self._h, C.byref(f), int(reset)), "j3dg_ctx_status")
        return f.value

    def set_profiling(self, on: bool):
        self._check(self._L.j3dg_ctx_set_profiling(self._h, int(on)), "j3dg_ctx_set_profiling")

    def timings(self, reset: bool = False) -> Timings:
        t = Timings()
        self._check(self._L.j3dg_ctx_timings(self._h, C.byref(t), int(reset)), "j3dg_ctx_timings")
        return t

    def set_matcap(self, matcap: np.ndarray, cavity: int):
        h, w = matcap.shape
        self._check(self._L.j3dg_ctx_set_matcap(self._h, _ptr(matcap), w, h, w, cavity), "j3dg_ctx_set_matcap")

    # -- meshes ---------------------------------------------------------------------
    def mesh_create(self, verts, tris, vcolors=None, uv=None, texture=None, cs=None, db_id: int = 0x20000000,
                    nv: int | None = None, nt: int | None = None) -> "Mesh":
        nv = verts.shape[0] if nv is None else nv
        nt = tris.shape[0] if nt is None else nt
        tw = th = 0
        if texture is not None:
            th, tw = texture.shape
        h = _vp()
        csb = None if cs is None else np.ascontiguousarray(cs, np.float32)
        self._check(self._L.j3dg_mesh_create(self._h, _ptr(verts), nv, _ptr(tris), nt, _ptr(vcolors), _ptr(uv),
                                             _ptr(texture), tw, th, tw, _ptr(csb), db_id, C.byref(h)), "j3dg_mesh_create")
        return Mesh(self, h)

    def mesh_create_empty(self, nv: int, nt: int, nr_nodes: int, cs=None, db_id: int = 0x20000000) -> "Mesh":
        h = _vp()
        csb = None if cs is None else np.ascontiguousarray(cs, np.float32)
        self._check(self._L.j3dg_mesh_create_empty(self._h, nv, nt, nr_nodes, _ptr(csb), db_id, C.byref(h)), "j3dg_mesh_create_empty")
        return Mesh(self, h)

    def cloud_create(self, pos, nrm=None, clr=None, cs=None, db_id: int = 0x40000000, n: int | None = None) -> "Cloud":
        n = pos.shape[0] if n is None else n
        h = _vp()
        csb = None if cs is None else np.ascontiguousarray(cs, np.float32)
        self._check(self._L.j3dg_cloud_create(self._h, _ptr(pos), _ptr(nrm), _ptr(clr), n, _ptr(csb), db_id, C.byref(h)), "j3dg_cloud_create")
        return Cloud(self, h, i_n=n)

    def ply_decode(self, data) -> "Ply":
        """jtk::read_ply for binary files: `data` = the whole file (bytes / uint8 array); decoded on the device."""
        buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
        h = _vp()
        self._check(self._L.j3dg_ply_decode(self._h, _ptr(buf), buf.size, C.byref(h)), "j3dg_ply_decode")
        return Ply(self, h)

    # -- frame stages -----------------------------------------------------------------
    @staticmethod
    def _handles(objs):
        arr = (_vp * max(1, len(objs)))(*[o._h for o in objs])
        return arr

    def cast(self, meshes, view: View, out=None, rect=None, stride: int | None = None):
        """canvas::update_canvas.  out: numpy PIXEL_DTYPE [H,W] (host) or device pointer/tensor."""
        w, h = view.width, view.height
        if out is None:
            out = np.zeros((h, w), PIXEL_DTYPE)
            out["object_id"] = 0xFFFFFFFF
        x0, y0, x1, y1 = rect if rect is not None else (0, 0, w - 1, h - 1)
        self._check(self._L.j3dg_cast(self._h, self._handles(meshes), len(meshes), C.byref(view), x0, y0, x1, y1,
                                      _ptr(out), stride or w), "j3dg_cast")
        return out

    def shade(self, pixels, view: View, matcap: np.ndarray, cavity: int, background=None, out=None,
              pixel_stride: int | None = None, rgba_stride: int | None = None):
        w, h = view.width, view.height
        if out is None:
            out = np.zeros((h, w), np.uint32)
        mh, mw = matcap.shape
        self._check(self._L.j3dg_shade(self._h, _ptr(pixels), pixel_stride or w, C.byref(view), _ptr(matcap), mw, mh, mw, cavity,
                                       _ptr(background), _ptr(out), rgba_stride or w), "j3dg_shade")
        return out

    def splat(self, clouds, view: View, pixels_in, pixels_inout, rgba_inout, pixel_stride: int | None = None,
              rgba_stride: int | None = None):
        w = view.width
        self._check(self._L.j3dg_splat(self._h, self._handles(clouds), len(clouds), C.byref(view), _ptr(pixels_in),
                                       _ptr(pixels_inout), pixel_stride or w, _ptr(rgba_inout), rgba_stride or w), "j3dg_splat")

    def render_frame(self, meshes, clouds, view: View, matcap=None, cavity: int = 0, bg_top=0xFF000000, bg_bottom=0xFF404040,
                     pixels_out=None, rgba_out=None):
        """view::render_scene in one call; matcap None = the one set by set_matcap."""
        mw = mh = 0
        if matcap is not None:
            mh, mw = matcap.shape
        self._check(self._L.j3dg_render_frame(self._h, self._handles(meshes), len(meshes), self._handles(clouds), len(clouds),
                                              C.byref(view), _ptr(matcap), mw, mh, mw, cavity, bg_top, bg_bottom,
                                              _ptr(pixels_out), _ptr(rgba_out)), "j3dg_render_frame")

    def frame_submit(self, meshes, clouds, view: View, matcap=None, cavity: int = 0, bg_top=0xFF000000, bg_bottom=0xFF404040,
                     pixels_out=None, rgba_out=None):
        """Pipelined render_frame: returns at once; the host buffers are complete after the matching frame_wait()."""
        mw = mh = 0
        if matcap is not None:
            mh, mw = matcap.shape
        self._check(self._L.j3dg_frame_submit(self._h, self._handles(meshes), len(meshes), self._handles(clouds), len(clouds),
                                              C.byref(view), _ptr(matcap), mw, mh, mw, cavity, bg_top, bg_bottom,
                                              _ptr(pixels_out), _ptr(rgba_out)), "j3dg_frame_submit")

    def frame_wait(self):
        self._check(self._L.j3dg_frame_wait(self._h), "j3dg_frame_wait")

    def set_dirty_rect(self, on: bool):
        """Host output buffers are persistent per-canvas buffers: copy only the rectangle that changed (include/j3dg.h)."""
        self._check(self._L.j3dg_ctx_set_dirty_rect(self._h, int(on)), "j3dg_ctx_set_dirty_rect")

    def readback_bytes(self, reset: bool = False) -> int:
        b = C.c_uint64()
        self._check(self._L.j3dg_ctx_readback_bytes(self._h, C.byref(b), int(reset)), "j3dg_ctx_readback_bytes")
        return b.value

    def set_tuning(self, lane_budget: int = 0, cast_algo: int = 0):
        self._check(self._L.j3dg_ctx_set_tuning(self._h, lane_budget, cast_algo), "j3dg_ctx_set_tuning")

    def set_screen_shard(self, rank: int = 0, world: int = 1):
        """Band b (J3DG_SHARD_BAND_ROWS = 32 rows) of every frame belongs to rank b mod world; world = 1: off."""
        self._check(self._L.j3dg_ctx_set_screen_shard(self._h, rank, world), "j3dg_ctx_set_screen_shard")

    def cast_stats(self, meshes, view: View):
        a, b = C.c_double(), C.c_double()
        self._check(self._L.j3dg_cast_stats(self._h, self._handles(meshes), len(meshes), C.byref(view), C.byref(a), C.byref(b)),
                    "j3dg_cast_stats")
        return a.value, b.value

    def pick(self, meshes, clouds, view: View, xy, pixels=None, pixel_stride: int | None = None) -> np.ndarray:
        """Device-side picking (canvas::get_pixel, view::get_id / get_world_position / get_index, pivot pick) for n
        query pixels xy [n,2] int32; pixels None = the resident canvas of the last frame.  Returns PICK_DTYPE [n]."""
        xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
        out = np.zeros((xy.shape[0],), PICK_DTYPE)
        self._check(self._L.j3dg_pick(self._h, self._handles(meshes), len(meshes), self._handles(clouds), len(clouds), C.byref(view),
                                      _ptr(pixels), pixel_stride or 0, _ptr(xy), xy.shape[0], _ptr(out)), "j3dg_pick")
        return out

    # -- result exchange over NVLink peer memory (include/j3dg.h; protocol: j3d_b200/dist.py::PeerFrames) ----------
    def peer_alloc(self, nbytes: int):
        """Zeroed device buffer + its CUDA-IPC handle (bytes).  Returns (device pointer, handle)."""
        p = _vp()
        h = (C.c_ubyte * 64)()
        self._check(self._L.j3dg_peer_alloc(self._h, nbytes, C.byref(p), C.cast(h, _vp)), "j3dg_peer_alloc")
        return p.value, bytes(h)

    def peer_free(self, ptr: int):
        self._check(self._L.j3dg_peer_free(self._h, _vp(ptr)), "j3dg_peer_free")

    def peer_open(self, handle: bytes) -> int:
        p = _vp()
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._check(self._L.j3dg_peer_open(self._h, C.cast(h, _vp), C.byref(p)), "j3dg_peer_open")
        return p.value

    def peer_close(self, ptr: int):
        self._check(self._L.j3dg_peer_close(self._h, _vp(ptr)), "j3dg_peer_close")

    def stream_signal(self, flag_ptr: int, value: int):
        self._check(self._L.j3dg_stream_signal(self._h, _vp(flag_ptr), value), "j3dg_stream_signal")

    def stream_wait_geq(self, flags_ptr: int, n: int, value: int):
        self._check(self._L.j3dg_stream_wait_geq(self._h, _vp(flags_ptr), n, value), "j3dg_stream_wait_geq")

    def stream_wait_timed_out(self) -> bool:
        t = C.c_int()
        self._check(self._L.j3dg_stream_wait_status(self._h, C.byref(t)), "j3dg_stream_wait_status")
        return bool(t.value)

    def cast_cost_image(self, meshes, view: View):
        """Per-pixel (node visits, triangle tests) of the counting pass — diagnostic."""
        n = np.zeros((view.height, view.width), np.uint32)
        t = np.zeros((view.height, view.width), np.uint32)
        self._check(self._L.j3dg_cast_cost_image(self._h, self._handles(meshes), len(meshes), C.byref(view), _ptr(n), _ptr(t)),
                    "j3dg_cast_cost_image")
        return n, t


class Mesh:
    def __init__(self, ctx: Context, h):
        self.ctx, self._h = ctx, h

    def destroy(self):
        if self._h:
            self.ctx._L.j3dg_mesh_destroy(self._h)
            self._h = _vp()

    def info(self) -> MeshInfo:
        i = MeshInfo()
        self.ctx._check(self.ctx._L.j3dg_mesh_info_get(self._h, C.byref(i)), "j3dg_mesh_info_get")
        return i

    def rebuild(self):
        self.ctx._check(self.ctx._L.j3dg_mesh_rebuild(self._h), "j3dg_mesh_rebuild")

    def set_cs(self, cs):
        csb = np.ascontiguousarray(cs, np.float32)
        self.ctx._check(self.ctx._L.j3dg_mesh_set_cs(self._h, _ptr(csb)), "j3dg_mesh_set_cs")

    def bvh_buffer(self, kind: int):
        p, n = _vp(), C.c_size_t()
        self.ctx._check(self.ctx._L.j3dg_mesh_bvh_buffer(self._h, kind, C.byref(p), C.byref(n)), "j3dg_mesh_bvh_buffer")
        return p.value, n.value

    def find_closest(self, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        hits = np.zeros((n, 4), np.float32)
        ids = np.zeros((n,), np.uint32)
        self.ctx._check(self.ctx._L.j3dg_mesh_find_closest(self._h, _ptr(rays), n, _ptr(hits), _ptr(ids)), "j3dg_mesh_find_closest")
        return hits, ids


    def find_all(self, rays: np.ndarray):
        """qbvh::find_all_triangles for a ray batch: (offsets [n+1], hits [total,4] {u,v,t,0}, triangle ids [total])."""
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        offsets = np.zeros((n + 1,), np.uint32)
        total = C.c_uint32()
        L = self.ctx._L
        self.ctx._check(L.j3dg_mesh_find_all(self._h, _ptr(rays), n, _ptr(offsets), None, None, 0, C.byref(total)), "j3dg_mesh_find_all")
        hits = np.zeros((max(total.value, 1), 4), np.float32)
        ids = np.zeros((max(total.value, 1),), np.uint32)
        self.ctx._check(L.j3dg_mesh_find_all(self._h, _ptr(rays), n, _ptr(offsets), _ptr(hits), _ptr(ids), total.value, C.byref(total)),
                        "j3dg_mesh_find_all")
        return offsets, hits[:total.value], ids[:total.value]

    def voxelize(self, max_dim: int, out=None) -> np.ndarray:
        """_write_vox's voxel grid (palette indices, 0 = empty) as [Z,Y,X] u8; out: optional device pointer / tensor."""
        dims = (C.c_uint32 * 3)()
        L = self.ctx._L
        self.ctx._check(L.j3dg_mesh_voxel_dims(self._h, max_dim, dims), "j3dg_mesh_voxel_dims")
        if out is None:
            out = np.zeros((dims[2], dims[1], dims[0]), np.uint8)
        cap = out.size if isinstance(out, np.ndarray) else (out.numel() if hasattr(out, "numel") else dims[0] * dims[1] * dims[2])
        self.ctx._check(L.j3dg_mesh_voxelize(self._h, max_dim, dims, _ptr(out), cap), "j3dg_mesh_voxelize")
        return out


class Ply:
    """j3dg_ply: the arrays of one decoded binary PLY file, resident on the device."""
    _KINDS = {"vertices": (0, np.float32, 3), "normals": (1, np.float32, 3), "colors": (2, np.uint32, 0),
              "triangles": (3, np.uint32, 3), "uv": (4, np.float32, 6)}

    def __init__(self, ctx: Context, h):
        self.ctx, self._h = ctx, h

    def destroy(self):
        if self._h:
            self.ctx._L.j3dg_ply_destroy(self._h)
            self._h = _vp()

    def info(self) -> PlyInfo:
        i = PlyInfo()
        self.ctx._check(self.ctx._L.j3dg_ply_info_get(self._h, C.byref(i)), "j3dg_ply_info_get")
        return i

    def array(self, kind: str) -> np.ndarray:
        which, dt, cols = self._KINDS[kind]
        i = self.info()
        rows = i.nr_of_faces if which >= 3 else i.nr_of_vertices
        if (which == 1 and not i.has_normals) or (which == 2 and not i.has_colors) or (which == 4 and not i.has_uv):
            rows = 0
        out = np.zeros((rows, cols) if cols else (rows,), dt)
        self.ctx._check(self.ctx._L.j3dg_ply_copy(self._h, which, _ptr(out), out.nbytes), "j3dg_ply_copy")
        return out

    def to_mesh(self, cs=None, db_id: int = 0x20000000) -> "Mesh":
        h = _vp()
        csb = None if cs is None else np.ascontiguousarray(cs, np.float32)
        self.ctx._check(self.ctx._L.j3dg_mesh_create_from_ply(self.ctx._h, self._h, _ptr(csb), db_id, C.byref(h)), "j3dg_mesh_create_from_ply")
        return Mesh(self.ctx, h)

    def to_cloud(self, cs=None, db_id: int = 0x40000000) -> "Cloud":
        h = _vp()
        csb = None if cs is None else np.ascontiguousarray(cs, np.float32)
        self.ctx._check(self.ctx._L.j3dg_cloud_create_from_ply(self.ctx._h, self._h, _ptr(csb), db_id, C.byref(h)), "j3dg_cloud_create_from_ply")
        return Cloud(self.ctx, h, i_n=self.info().nr_of_vertices)


class Cloud:
    def __init__(self, ctx: Context, h, i_n: int = 0):
        self.ctx, self._h, self.n = ctx, h, i_n

    def estimate_normals(self, k: int) -> np.ndarray:
        """estimate_normals (j3d/pc.cpp:256): k-NN + plane fit on the device, orientation propagation; replaces the cloud's normals."""
        out = np.zeros((self.n, 3), np.float32)
        self.ctx._check(self.ctx._L.j3dg_cloud_estimate_normals(self._h, k, _ptr(out)), "j3dg_cloud_estimate_normals")
        return out

    def knn_normals(self, k: int):
        """Unoriented normals + the neighbour lists (ascending distance)."""
        nrm = np.zeros((self.n, 3), np.float32)
        knn = np.zeros((self.n, min(k, self.n)), np.uint32)
        self.ctx._check(self.ctx._L.j3dg_cloud_knn_normals(self._h, k, _ptr(nrm), _ptr(knn)), "j3dg_cloud_knn_normals")
        return nrm, knn

    def destroy(self):
        if self._h:
            self.ctx._L.j3dg_cloud_destroy(self._h)
            self._h = _vp()


# ---------------------------------------------------------------------------------------
# multi-GPU: one process per GPU (include/j3dg.h "multi-GPU"; csrc/group.cu)
# ---------------------------------------------------------------------------------------
def group_unique_id() -> bytes:
    """The NCCL id one rank creates and ships to the others (any channel) before Group(...)."""
    buf = (C.c_ubyte * 128)()
    rc = lib().j3dg_group_unique_id(C.cast(buf, _vp))
    if rc != 0:
        raise J3dgError(f"j3dg_group_unique_id = {rc}: {lib().j3dg_last_error(None).decode()}")
    return bytes(buf)


class Group:
    """j3dg_group: this context's membership in a set of `world` processes, one per GPU.  Every call is collective."""

    def __init__(self, ctx: Context, rank: int, world: int, unique_id: bytes):
        self.ctx, self.rank, self.world = ctx, rank, world
        self._h = _vp()
        idb = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        ctx._check(ctx._L.j3dg_group_create(ctx._h, rank, world, C.cast(idb, _vp), C.byref(self._h)), "j3dg_group_create")

    def destroy(self):
        if self._h:
            self.ctx._L.j3dg_group_destroy(self._h)
            self._h = _vp()

    def barrier(self):
        self.ctx._check(self.ctx._L.j3dg_group_barrier(self._h), "j3dg_group_barrier")

    def max_float(self, *values: float) -> list[float]:
        arr = (C.c_float * len(values))(*values)
        self.ctx._check(self.ctx._L.j3dg_group_max_float(self._h, C.cast(arr, _vp), len(values)), "j3dg_group_max_float")
        return list(arr)

    def allreduce_max_u64(self, dev_ptr: int, n: int):
        self.ctx._check(self.ctx._L.j3dg_group_allreduce_max_u64(self._h, _vp(dev_ptr), n), "j3dg_group_allreduce_max_u64")

    def broadcast_mesh(self, mesh: "Mesh | None", root: int = 0) -> "Mesh":
        """root passes its built mesh, the other ranks None; every rank returns a mesh holding the same BVH + geometry."""
        h = mesh._h if mesh is not None else _vp()
        self.ctx._check(self.ctx._L.j3dg_group_broadcast_mesh(self._h, root, C.byref(h)), "j3dg_group_broadcast_mesh")
        return mesh if mesh is not None else Mesh(self.ctx, h)

    def frames(self, width: int, height: int, dst: int = 0, shared_frame: bool = False, nslots: int = 2) -> "Frames":
        return Frames(self, width, height, dst, shared_frame, nslots)


class Frames:
    """j3dg_frames: every rank renders straight into rank dst's HBM (include/j3dg.h has the protocol)."""

    def __init__(self, group: Group, width: int, height: int, dst: int, shared_frame: bool, nslots: int = 2):
        self.group, self.ctx, self.w, self.h, self.dst, self.shared, self.nslots = group, group.ctx, width, height, dst, shared_frame, nslots
        self._h = _vp()
        self.ctx._check(self.ctx._L.j3dg_frames_create_n(group._h, width, height, dst, int(shared_frame), nslots, C.byref(self._h)), "j3dg_frames_create_n")

    def destroy(self):
        if self._h:
            self.ctx._L.j3dg_frames_destroy(self._h)
            self._h = _vp()

    def set_lane(self, slot: int, ctx: "Context"):
        self.ctx._check(self.ctx._L.j3dg_frames_set_lane(self._h, slot, ctx._h), "j3dg_frames_set_lane")

    def begin(self) -> int:
        k = _u32()
        self.ctx._check(self.ctx._L.j3dg_frames_begin(self._h, C.byref(k)), "j3dg_frames_begin")
        return k.value

    def target(self, k: int) -> int:
        p = _vp()
        self.ctx._check(self.ctx._L.j3dg_frames_target(self._h, k, C.byref(p)), "j3dg_frames_target")
        return p.value

    def arrive(self, k: int):
        self.ctx._check(self.ctx._L.j3dg_frames_arrive(self._h, k), "j3dg_frames_arrive")

    def release(self, k: int):
        self.ctx._check(self.ctx._L.j3dg_frames_release(self._h, k), "j3dg_frames_release")

    def view(self, k: int) -> int:
        p = _vp()
        self.ctx._check(self.ctx._L.j3dg_frames_view(self._h, k, C.byref(p)), "j3dg_frames_view")
        return p.value
