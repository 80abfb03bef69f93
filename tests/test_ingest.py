"""Ingest: binary PLY decode and point-cloud normal estimation.
CPU tests pin the numpy restatement (oracle/ingest_oracle.py) on the unmodified reference — live through
oracle/_ref/libj3d_ref.so where it exists, and through tests/golden/ingest.npz (made from it) everywhere.
GPU tests compare the CUDA path (j3dg_ply_decode, j3dg_cloud_knn_normals, j3dg_cloud_estimate_normals through the C ABI)
with that oracle, the goldens and the live reference.  Integer / byte outputs: bit-exact.  Normals: the reference runs a
float SVD on the float scatter matrix, the CUDA path a double Jacobi on the SAME matrix; they agree to the float SVD's
accuracy, |n_gpu . n_ref| >= 1 - 1e-4 wherever the two smallest eigenvalues are separated (gap ratio > 0.05)."""
import sys
from pathlib import Path

import numpy as np
import pytest

import j3d_b200 as j
from ply_util import sample_files, write_ply

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "golden"))
from make_golden_ingest import NORMALS_K, normals_cloud  # noqa: E402
from oracle import ingest_oracle as io  # noqa: E402
from oracle.bindings import ref_available  # noqa: E402

GOLD = np.load(HERE / "golden" / "ingest.npz")
KINDS = ("vertices", "normals", "colors", "triangles", "uv")
NORMAL_TOL = 1e-4


def _same(a, b):
    return a.shape == b.shape and a.tobytes() == b.tobytes()


# ---- CPU: the oracle against the reference -------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(sample_files().keys()))
def test_oracle_ply_matches_golden(name):
    got = io.read_ply_binary(sample_files()[name])
    for k in KINDS:
        assert _same(got[k], GOLD[f"{name}/{k}"]), (name, k)


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
def test_oracle_ply_equals_reference_live():
    from oracle.bindings import ref_read_ply
    rng = np.random.default_rng(77)
    for trial in range(6):
        nv, nf = int(rng.integers(3, 400)), int(rng.integers(1, 300))
        v = rng.normal(size=(nv, 3)) * 10.0 ** int(rng.integers(-3, 4))
        faces = [list(rng.integers(0, nv, size=int(rng.integers(3, 7)))) for _ in range(nf)]
        data = write_ply([("x", "f4", v[:, 0]), ("y", "f8", v[:, 1]), ("z", "f4", v[:, 2]), ("red", "u1", rng.integers(0, 256, nv)), ("blue", "u1", rng.integers(0, 256, nv))],
                         faces, big_endian=bool(trial & 1), index_type=("i4", "u4", "u2")[trial % 3] if nv < 60000 else "i4")
        want = ref_read_ply(data)
        got = io.read_ply_binary(data)
        for k in KINDS:
            assert _same(got[k], want[k]), (trial, k)


def _check_normals(got, want, gap, oriented):
    d = np.einsum("ij,ij->i", got.astype(np.float64), want.astype(np.float64))
    ok = gap > 0.05
    assert ok.mean() > 0.95
    assert (np.abs(d[ok]) >= 1.0 - NORMAL_TOL).all(), float(np.abs(d[ok]).min())
    if oriented:  # one connected surface: the sign of the whole component follows the seed's SVD sign
        agree = (d[ok] > 0).mean()
        assert max(agree, 1.0 - agree) >= 0.995, agree


def test_oracle_normals_match_golden():
    pos = normals_cloud()
    nb = io.knn(pos, NORMALS_K)
    nrm, gap = io.fit_normals(pos, nb)
    _check_normals(io.orient(nrm, nb), GOLD["normals/k10"], gap, oriented=True)


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
def test_oracle_normals_equal_reference_live():
    from oracle.bindings import ref_estimate_normals
    pos = normals_cloud(1500, seed=5) * np.float32(37.5) + np.float32(100.0)
    nb = io.knn(pos, 16)
    nrm, gap = io.fit_normals(pos, nb)
    _check_normals(io.orient(nrm, nb), ref_estimate_normals(pos, 16), gap, oriented=True)


# ---- GPU: the CUDA path against the oracle, the goldens and the live reference ----------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(sample_files().keys()))
def test_ply_decode_matches_golden(ctx, name):
    data = sample_files()[name]
    ply = ctx.ply_decode(data)
    info = ply.info()
    assert info.file_bytes == len(data) and info.format == (2 if b"binary_big_endian" in data[:64] else 1)
    for k in KINDS:
        assert _same(ply.array(k), GOLD[f"{name}/{k}"]), (name, k)
    ply.destroy()


@pytest.mark.gpu
def test_ply_decode_large_random_files_equal_oracle(ctx):
    """300 k vertices / 600 k faces, odd record sizes, both byte orders, one file with mixed polygons (the offset path)."""
    rng = np.random.default_rng(9)
    verts, tris = j.icosphere(173)
    nv = verts.shape[0]
    col = rng.integers(0, 256, size=(nv, 3))
    for be in (False, True):
        props = [("x", "f4", verts[:, 0]), ("confidence", "u1", rng.integers(0, 255, nv)), ("y", "f4", verts[:, 1]), ("z", "f4", verts[:, 2]),
                 ("red", "u1", col[:, 0]), ("green", "u1", col[:, 1]), ("blue", "u1", col[:, 2])]
        faces = tris.tolist()
        if be:
            faces = [f + [f[1]] if i % 1000 == 7 else f for i, f in enumerate(faces)]
        data = write_ply(props, faces, big_endian=be)
        want = io.read_ply_binary(data) if not be else None
        ply = ctx.ply_decode(data)
        assert _same(ply.array("vertices"), verts) and _same(ply.array("triangles"), tris)
        c = ply.array("colors")
        assert (c == (col[:, 0] | (col[:, 1] << 8) | (col[:, 2] << 16) | 0xFF000000).astype(np.uint32)).all()
        if want is not None:
            for k in KINDS:
                assert _same(ply.array(k), want[k]), k
        ply.destroy()


@pytest.mark.gpu
def test_ply_decode_rejects_bad_files(ctx):
    good = sample_files()["mesh_le"]
    with pytest.raises(j.J3dgError):
        ctx.ply_decode(good.replace(b"binary_little_endian", b"ascii"))
    with pytest.raises(j.J3dgError):
        ctx.ply_decode(good[: len(good) // 2])
    with pytest.raises(j.J3dgError):
        ctx.ply_decode(b"plx\nformat binary_little_endian 1.0\nend_header\n")
    with pytest.raises(j.J3dgError):
        ctx.ply_decode(good[:40])
    verts, tris = j.icosphere(2)
    two = write_ply([("x", "f4", verts[:, 0]), ("y", "f4", verts[:, 1]), ("z", "f4", verts[:, 2])], [[0, 1, 2], [0, 1]])
    with pytest.raises(j.J3dgError):
        ctx.ply_decode(two)  # a face with fewer than 3 vertices


@pytest.mark.gpu
def test_mesh_and_cloud_from_ply_render_like_arrays(ctx):
    """j3dg_mesh_create_from_ply / j3dg_cloud_create_from_ply: the frame equals the one of the same arrays uploaded directly."""
    rng = np.random.default_rng(4)
    verts, tris = j.icosphere(24)
    col = rng.integers(0, 256, size=(verts.shape[0], 3))
    data = write_ply([("x", "f4", verts[:, 0]), ("y", "f4", verts[:, 1]), ("z", "f4", verts[:, 2]),
                      ("red", "u1", col[:, 0]), ("green", "u1", col[:, 1]), ("blue", "u1", col[:, 2])], tris.tolist())
    ply = ctx.ply_decode(data)
    m1 = ply.to_mesh()
    vc = (col.astype(np.float32) / np.float32(255.0)).astype(np.float32)
    m2 = ctx.mesh_create(verts, tris, vcolors=vc)
    mn, mx = j.compute_bb(verts)
    view = j.orbit_view(j.make_view(320, 200, mn, mx, j.DEFAULT_FLAGS | j.VERTEXCOLORS), 25.0)
    mc, cav = j.make_matcap(0)
    out = []
    for m in (m1, m2):
        px = np.zeros((200, 320), j.PIXEL_DTYPE)
        rgba = np.zeros((200, 320), np.uint32)
        ctx.render_frame([m], [], view, mc, cav, pixels_out=px, rgba_out=rgba)
        out.append((px.copy(), rgba.copy()))
    assert out[0][0].tobytes() == out[1][0].tobytes() and out[0][1].tobytes() == out[1][1].tobytes()
    assert (out[0][0]["object_id"] != 0xFFFFFFFF).mean() > 0.2
    pos, nrm, clr = j.cloud(5000)
    cdata = write_ply([("x", "f4", pos[:, 0]), ("y", "f4", pos[:, 1]), ("z", "f4", pos[:, 2]), ("nx", "f4", nrm[:, 0]), ("ny", "f4", nrm[:, 1]), ("nz", "f4", nrm[:, 2]),
                       ("red", "u1", clr & 255), ("green", "u1", (clr >> 8) & 255), ("blue", "u1", (clr >> 16) & 255), ("alpha", "u1", (clr >> 24) & 255)])
    cply = ctx.ply_decode(cdata)
    c1 = cply.to_cloud()
    c2 = ctx.cloud_create(pos, nrm, clr)
    outs = []
    for c in (c1, c2):
        px = np.zeros((200, 320), j.PIXEL_DTYPE)
        rgba = np.zeros((200, 320), np.uint32)
        ctx.render_frame([], [c], view, mc, cav, pixels_out=px, rgba_out=rgba)
        outs.append((px.copy(), rgba.copy()))
    assert outs[0][0].tobytes() == outs[1][0].tobytes() and outs[0][1].tobytes() == outs[1][1].tobytes()
    for o in (m1, m2, c1, c2, ply, cply):
        o.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 6, 10, 33, 64])
def test_knn_equals_bruteforce(ctx, k):
    pos = normals_cloud(2500, seed=21 + k)
    if k == 6:
        pos = (pos * np.float32(250.0) + np.float32([1000.0, -300.0, 50.0])).astype(np.float32)  # far from the origin
    c = ctx.cloud_create(pos)
    nrm, nb = c.knn_normals(k)
    want = io.knn(pos, k)
    assert (nb == want).all()
    wn, gap = io.fit_normals(pos, want)
    if k >= 6:
        d = np.abs(np.einsum("ij,ij->i", nrm.astype(np.float64), wn.astype(np.float64)))
        assert (d[gap > 1e-3] >= 1.0 - 1e-6).all()
    c.destroy()


@pytest.mark.gpu
def test_knn_degenerate_clouds(ctx):
    """Fewer points than k, all points equal, points on a line / in a plane, heavy duplicates."""
    rng = np.random.default_rng(2)
    few = rng.normal(size=(5, 3)).astype(np.float32)
    c = ctx.cloud_create(few)
    nrm, nb = c.knn_normals(10)
    assert nb.shape == (5, 5) and (np.sort(nb, axis=1) == np.arange(5)).all()
    c.destroy()
    same = np.ones((300, 3), np.float32)
    c = ctx.cloud_create(same)
    nrm, nb = c.knn_normals(8)
    assert (nb == io.knn(same, 8)).all() and np.isfinite(nrm).all()
    c.destroy()
    plane = rng.random(size=(3999, 3)).astype(np.float32)
    plane[:, 2] = 0.25
    plane[::3] = plane[1::3][: plane[::3].shape[0]]  # duplicates
    c = ctx.cloud_create(plane)
    nrm, nb = c.knn_normals(12)
    assert (nb == io.knn(plane, 12)).all()
    assert (np.abs(nrm[:, 2]) > 0.999).all()
    c.destroy()


@pytest.mark.gpu
def test_estimate_normals_matches_golden(ctx):
    pos = normals_cloud()
    c = ctx.cloud_create(pos)
    got = c.estimate_normals(NORMALS_K)
    _, gap = io.fit_normals(pos, io.knn(pos, NORMALS_K))
    _check_normals(got, GOLD["normals/k10"], gap, oriented=True)
    # and exactly the oracle's orientation walk on the CUDA path's own unoriented normals
    nrm, nb = c.knn_normals(NORMALS_K)
    assert (io.orient(nrm, nb) == got).all()
    c.destroy()


@pytest.mark.gpu
@pytest.mark.skipif(not ref_available(), reason="oracle/_ref did not travel")
def test_estimate_normals_equals_reference_live_200k(ctx):
    from oracle.bindings import ref_estimate_normals
    pos = normals_cloud(200_000, seed=8)
    c = ctx.cloud_create(pos)
    got = c.estimate_normals(12)
    want = ref_estimate_normals(pos, 12)
    d = np.einsum("ij,ij->i", got.astype(np.float64), want.astype(np.float64))
    good = np.abs(d) >= 1.0 - NORMAL_TOL
    assert good.mean() > 0.97, good.mean()   # the rest: near-degenerate neighbourhoods (two close eigenvalues)
    agree = (d[good] > 0).mean()
    assert max(agree, 1.0 - agree) > 0.99, agree
    # the cloud now shades with the estimated normals
    mn, mx = j.compute_bb(pos)
    view = j.make_view(320, 200, mn, mx, j.DEFAULT_FLAGS)
    mc, cav = j.make_matcap(0)
    rgba = np.zeros((200, 320), np.uint32)
    ctx.render_frame([], [c], view, mc, cav, rgba_out=rgba)
    assert len(np.unique(rgba)) > 50
    c.destroy()


def test_bench_ply_writer_is_read_by_the_oracle_and_the_reference():
    """bench.py's in-memory PLY writer (the ingest stage of the default line) produces a file that the oracle — and the
    unmodified jtk reader where it is available — decode back to the same arrays."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", str(HERE.parent / "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    verts, tris = j.icosphere(7)
    data = bench.ply_bytes(verts, tris)
    got = io.read_ply_binary(data)
    assert _same(got["vertices"], verts) and _same(got["triangles"], tris)
    if ref_available():
        from oracle.bindings import ref_read_ply
        want = ref_read_ply(data)
        assert _same(want["vertices"], verts) and _same(want["triangles"], tris)
