"""Writes binary PLY files in memory for the ingest tests (any scalar type per property, either byte order, optional
normals / colours / extra properties, mixed polygons, texcoord lists)."""
from __future__ import annotations

import numpy as np

_NAMES = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float", "f8": "double"}


def write_ply(vertex_props, faces=None, big_endian=False, index_type="i4", count_type="u1", index_name="vertex_indices",
              texcoords=None, face_extra=None, extra_element=None, comment=True) -> bytes:
    """vertex_props: list of (name, dtype code, 1-D array); faces: list of index lists; texcoords: list of float lists
    (one per face) or None; face_extra: (name, dtype code, array) scalar property written BEFORE the lists;
    extra_element: (name, count, dtype code) element of scalars written between vertices and faces."""
    order = ">" if big_endian else "<"
    nv = len(vertex_props[0][2]) if vertex_props else 0
    head = ["ply", f"format {'binary_big_endian' if big_endian else 'binary_little_endian'} 1.0"]
    if comment:
        head.append("comment written by tests/ply_util.py")
    head.append(f"element vertex {nv}")
    for name, ty, _ in vertex_props:
        head.append(f"property {_NAMES[ty]} {name}")
    if extra_element:
        head.append(f"element {extra_element[0]} {extra_element[1]}")
        head.append(f"property {_NAMES[extra_element[2]]} value")
    if faces is not None:
        head.append(f"element face {len(faces)}")
        if face_extra:
            head.append(f"property {_NAMES[face_extra[1]]} {face_extra[0]}")
        head.append(f"property list {_NAMES[count_type]} {_NAMES[index_type]} {index_name}")
        if texcoords is not None:
            head.append("property list uchar float texcoord")
    head.append("end_header")
    out = bytearray(("\n".join(head) + "\n").encode("ascii"))
    if nv:
        dt = np.dtype([(name, order + ty) for name, ty, _ in vertex_props])
        rec = np.zeros(nv, dt)
        for name, _, arr in vertex_props:
            rec[name] = arr
        out += rec.tobytes()
    if extra_element:
        out += np.arange(extra_element[1]).astype(order + extra_element[2]).tobytes()
    if faces is not None:
        for f, idx in enumerate(faces):
            if face_extra:
                out += np.array([face_extra[2][f]]).astype(order + face_extra[1]).tobytes()
            out += np.array([len(idx)]).astype(order + count_type).tobytes()
            out += np.asarray(idx).astype(order + index_type).tobytes()
            if texcoords is not None:
                out += np.array([len(texcoords[f])]).astype("u1").tobytes()
                out += np.asarray(texcoords[f], np.float64).astype(order + "f4").tobytes()
    return bytes(out)


def sample_files(seed=3):
    """name -> bytes: the cases the ingest tests decode."""
    import j3d_b200 as j
    rng = np.random.default_rng(seed)
    verts, tris = j.icosphere(6)
    nv, nt = verts.shape[0], tris.shape[0]
    nrm = verts / np.linalg.norm(verts, axis=1, keepdims=True)
    col = rng.integers(0, 256, size=(nv, 4), dtype=np.uint8)
    xyz = [("x", "f4", verts[:, 0]), ("y", "f4", verts[:, 1]), ("z", "f4", verts[:, 2])]
    nxyz = [("nx", "f4", nrm[:, 0]), ("ny", "f4", nrm[:, 1]), ("nz", "f4", nrm[:, 2])]
    rgb = [("red", "u1", col[:, 0]), ("green", "u1", col[:, 1]), ("blue", "u1", col[:, 2])]
    files = {}
    files["mesh_le"] = write_ply(xyz, tris.tolist())
    files["mesh_be_colors_normals"] = write_ply(xyz + nxyz + rgb + [("alpha", "u1", col[:, 3])], tris.tolist(), big_endian=True)
    files["cloud_only"] = write_ply(xyz + nxyz + rgb)
    files["double_coords_short_colors"] = write_ply(
        [("x", "f8", verts[:, 0].astype(np.float64) * 1.000000123), ("y", "f8", verts[:, 1]), ("quality", "f4", rng.random(nv)), ("z", "f8", verts[:, 2]),
         ("r", "u2", col[:, 0]), ("g", "i2", col[:, 1]), ("b", "i4", col[:, 2])], tris.tolist(), index_type="u4", count_type="i4", index_name="vertex_index")
    files["diffuse_colors_be"] = write_ply(xyz + [("diffuse_red", "u1", col[:, 0]), ("diffuse_green", "u1", col[:, 1]), ("diffuse_blue", "u1", col[:, 2])],
                                           tris.tolist(), big_endian=True, index_type="u2")
    quads = [list(t) + [int(t[0])] if i % 3 == 0 else list(t) for i, t in enumerate(tris.tolist())]
    files["mixed_polygons"] = write_ply(xyz, quads)
    uvs = [list(rng.random(6)) for _ in range(nt)]
    files["texcoords"] = write_ply(xyz, tris.tolist(), texcoords=uvs)
    ragged = [list(rng.random(int(rng.integers(1, 9)))) for _ in range(nt)]
    files["texcoords_ragged_be"] = write_ply(xyz, tris.tolist(), texcoords=ragged, big_endian=True)
    files["face_scalar_and_extra_element"] = write_ply(xyz, tris.tolist(), face_extra=("flags", "u1", rng.integers(0, 255, nt)), extra_element=("edge", 17, "i4"))
    files["int_coords"] = write_ply([("x", "i2", (verts[:, 0] * 1000).astype(np.int16)), ("y", "i4", (verts[:, 1] * 1e6).astype(np.int32)), ("z", "i1", (verts[:, 2] * 100).astype(np.int8))], tris.tolist())
    files["no_faces_no_comment"] = write_ply(xyz, None, comment=False)
    return files
