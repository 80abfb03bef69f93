"""j3d_b200 — B200-native (sm_100a CUDA) implementation of j3d's render hot path:
GPU BVH build -> per-pixel ray cast -> fused shading -> point-cloud splat, behind the C ABI in
include/j3dg.h.  See DESIGN.md.  There is no CPU fallback: the CUDA library must be built
(`python -m j3d_b200.build`) and a B200 must be present to render.
"""
from .capi import (  # noqa: F401
    DEFAULT_FLAGS, EDGES, ONE_BIT, PICK_DTYPE, PIXEL_DTYPE, SHADING, SHADOW, TEXTURED, VERTEXCOLORS, WIREFRAME,
    Cloud, Context, Frames, Group, J3dgError, Mesh, MeshInfo, Ply, PlyInfo, Timings, View, group_unique_id,
    cloud, compute_bb, fill_background, icosphere, make_matcap, make_view, orbit_view, vertex_colors,
)

__all__ = [n for n in dir() if not n.startswith("_")]
