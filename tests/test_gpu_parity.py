"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes over
include/j3dg.h), against the CPU oracle (oracle/j3d_oracle.c) on the same seeded inputs, and
against the unmodified reference (oracle/_ref) when its prebuilt library travelled along.
Run with `pytest -m gpu` on a B200.
"""
import numpy as np
import pytest

import j3d_b200 as j
from parity import compare_pixels, compare_rgba, MISS

pytestmark = pytest.mark.gpu

FMAX = float(np.finfo(np.float32).max)


def cube():
    v = np.array([[-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1], [-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1]], np.float32)
    t = np.array([[0, 1, 2], [0, 2, 3], [7, 6, 5], [7, 5, 4], [1, 0, 4], [1, 4, 5], [2, 1, 5], [2, 5, 6], [3, 2, 6], [3, 6, 7], [0, 3, 7], [0, 7, 4]], np.uint32)
    return v, t


def test_cube_known_answer(ctx):
    """jtk.tests/qbvh_tests.cpp:708-749 (find_closest_with_ray) through j3dg_mesh_find_closest."""
    v, t = cube()
    m = ctx.mesh_create(v, t)
    rays = np.array([[0.8, -0.5, 0.8, 1, 0, 0, 0, FMAX], [0.8, -0.5, 0.8, 1, 0, 0, -FMAX, 0], [0.8, -0.5, 0.8, 1, 0, 0, -FMAX, FMAX]], np.float32)
    hits, ids = m.find_closest(rays)
    assert abs(hits[0, 2] - 0.2) <= 1e-6 and ids[0] == 6
    assert abs(hits[1, 2] + 1.8) <= 1e-6 and ids[1] == 10
    assert abs(hits[2, 2] - 0.2) <= 1e-6 and ids[2] == 6
    assert (hits[:, 3] == 1).all()
    m.destroy()


def test_find_closest_random_rays(ctx, oracle):
    verts, tris = j.icosphere(9)
    rng = np.random.default_rng(5)
    n = 4000
    org = rng.normal(size=(n, 3)).astype(np.float32) * 2.0
    d = rng.normal(size=(n, 3)).astype(np.float32)
    rays = np.concatenate([org, d, np.full((n, 1), -FMAX, np.float32), np.full((n, 1), FMAX, np.float32)], axis=1).astype(np.float32)
    rays[: n // 2, 6] = 0.0
    m = ctx.mesh_create(verts, tris)
    hits, ids = m.find_closest(rays)
    om = oracle.mesh(verts, tris)
    ohits, oids = om.find_closest(rays)
    assert (hits[:, 3] == ohits[:, 3]).mean() > 0.999
    both = (hits[:, 3] == 1) & (ohits[:, 3] == 1)
    assert (ids[both] == oids[both]).mean() > 0.999
    same = both & (ids == oids)
    assert np.abs(hits[same, 2] - ohits[same, 2]).max() <= 1e-5 * np.abs(ohits[same, 2]).max()
    m.destroy(); om.destroy()


def _scene(f, w, h, flags, angle, with_colors=False):
    verts, tris = j.icosphere(f)
    vc = j.vertex_colors(verts) if with_colors else None
    mn, mx = j.compute_bb(verts)
    v = j.make_view(w, h, mn, mx, flags)
    if angle:
        v = j.orbit_view(v, angle)
    return verts, tris, vc, v


@pytest.mark.parametrize("f,w,h,flags,angle,colors", [
    (12, 480, 270, j.DEFAULT_FLAGS, 0.0, False),
    (40, 640, 360, j.DEFAULT_FLAGS | j.SHADOW, 33.0, True),
    (25, 333, 201, j.DEFAULT_FLAGS, -71.0, True),            # ragged canvas size (not a multiple of the tile)
    (3, 64, 48, j.DEFAULT_FLAGS | j.SHADOW, 10.0, False),     # 180 triangles
])
def test_cast_matches_oracle(ctx, oracle, f, w, h, flags, angle, colors):
    verts, tris, vc, v = _scene(f, w, h, flags, angle, colors)
    m = ctx.mesh_create(verts, tris, vcolors=vc)
    got = ctx.cast([m], v)
    om = oracle.mesh(verts, tris, vcolors=vc)
    want = oracle.cast([om], v)
    st = compare_pixels(got, want, tag=f"f={f}")
    assert st["hits"] > 0
    m.destroy(); om.destroy()


@pytest.mark.parametrize("budget,algo", [(1, 0), (3, 0), (8, 0), (0, 1)])
def test_cast_independent_of_tuning(ctx, oracle, budget, algo):
    """Both traversal kernels and the hard-ray hand-over between them: a lane budget of 1 sends every ray
    that meets the mesh through the 8-lanes-per-ray kernel; algo 1 uses that kernel alone.  Shadow rays
    take the same two paths.  The result must match the oracle either way."""
    verts, tris, vc, v = _scene(40, 640, 360, j.DEFAULT_FLAGS | j.SHADOW, 33.0, True)
    m = ctx.mesh_create(verts, tris, vcolors=vc)
    try:
        ctx.set_tuning(1 << 30, 0)
        base = ctx.cast([m], v)
        ctx.set_tuning(budget, algo)
        got = ctx.cast([m], v)
    finally:
        ctx.set_tuning(0, 0)
    om = oracle.mesh(verts, tris, vcolors=vc)
    want = oracle.cast([om], v)
    st = compare_pixels(got, want, tag=f"budget={budget} algo={algo}")
    assert st["hits"] > 0
    # against the lane kernel alone: the same closest hits (exact ties may pick the other triangle)
    assert (got["object_id"] == base["object_id"]).mean() >= 0.9999
    same = got["object_id"] == base["object_id"]
    assert (got["depth"][same] == base["depth"][same]).all()
    assert (got["mark"][same] == base["mark"][same]).mean() >= 0.9999
    m.destroy(); om.destroy()


def test_cast_two_objects_textured(ctx, oracle):
    w, h = 640, 360
    verts, tris = j.icosphere(30)
    vc = j.vertex_colors(verts)
    v2, t2 = j.icosphere(10)
    cs = np.eye(4, dtype=np.float32)
    a = 0.7
    cs[0, 0], cs[0, 1], cs[1, 0], cs[1, 1] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
    cs[3, :3] = [1.5, 0.3, -0.5]
    uv = np.random.default_rng(1).random((t2.shape[0], 6), dtype=np.float32)
    tex = (np.random.default_rng(2).integers(0, 2 ** 32, (64, 48), dtype=np.uint64)).astype(np.uint32) | np.uint32(0xFF000000)
    mn = np.minimum(verts.min(0), v2.min(0) + cs[3, :3] - 0.2)
    mx = np.maximum(verts.max(0), v2.max(0) + cs[3, :3] + 0.2)
    v = j.orbit_view(j.make_view(w, h, mn, mx, j.DEFAULT_FLAGS | j.SHADOW), -20.0)
    m1 = ctx.mesh_create(verts, tris, vcolors=vc, db_id=0x20000000)
    m2 = ctx.mesh_create(v2, t2, uv=uv, texture=tex, cs=cs.reshape(-1), db_id=0x20000001)
    got = ctx.cast([m1, m2], v)
    o1 = oracle.mesh(verts, tris, vcolors=vc, db_id=0x20000000)
    o2 = oracle.mesh(v2, t2, uv=uv, texture=tex, cs=cs.reshape(-1), db_id=0x20000001)
    want = oracle.cast([o1, o2], v)
    compare_pixels(got, want, tag="two objects")
    assert (got["db_id"] == 0x20000001).sum() > 1000
    for x in (m1, m2):
        x.destroy()


def test_cast_rect_and_empty_scene(ctx, oracle):
    verts, tris, vc, v = _scene(10, 200, 120, j.DEFAULT_FLAGS, 5.0)
    m = ctx.mesh_create(verts, tris)
    full = ctx.cast([m], v)
    part = np.zeros_like(full)
    part["object_id"] = 0x12345678  # must stay untouched outside the rectangle
    ctx.cast([m], v, out=part, rect=(50, 30, 149, 89))
    assert (part[30:90, 50:150].tobytes() == full[30:90, 50:150].tobytes())
    outside = np.ones(part.shape, bool)
    outside[30:90, 50:150] = False
    assert (part["object_id"][outside] == 0x12345678).all()
    # clamping of an out-of-range rectangle (canvas.cpp:682-698)
    clamped = ctx.cast([m], v, rect=(-5, -5, 10_000, 10_000))
    assert clamped.tobytes() == full.tobytes()
    # no objects: every pixel is the miss record (canvas.cpp:744-760)
    empty = ctx.cast([], v)
    assert (empty["object_id"] == MISS).all() and (empty["db_id"] == 0).all() and (empty["depth"] == np.float32(FMAX)).all()
    m.destroy()


@pytest.mark.parametrize("flags", [
    j.DEFAULT_FLAGS, j.DEFAULT_FLAGS & ~j.EDGES, j.DEFAULT_FLAGS | j.WIREFRAME, j.DEFAULT_FLAGS | j.ONE_BIT,
    j.DEFAULT_FLAGS | j.SHADOW, j.SHADOW | j.EDGES | j.VERTEXCOLORS, j.SHADOW | j.VERTEXCOLORS,
])
def test_shade_matches_oracle(ctx, oracle, flags):
    """Shading in isolation: both sides shade the SAME (oracle-made) pixel buffer."""
    w, h = 512, 300
    verts, tris, vc, v = _scene(30, w, h, flags, 33.0, True)
    om = oracle.mesh(verts, tris, vcolors=vc)
    px = oracle.cast([om], v)
    mc, cav = j.make_matcap(0)
    bg = j.fill_background(w, h)
    want = oracle.shade(px, v, mc, cav, bg.copy())
    got = ctx.shade(px, v, mc, cav, background=bg)
    st = compare_rgba(got, want, tag=hex(flags))
    assert st["gt1"] <= 0.0005 * w * h
    # background == None keeps the caller's content on miss pixels
    mine = np.full((h, w), 0xFF123456, np.uint32)
    got2 = ctx.shade(px, v, mc, cav, background=None, out=mine)
    miss = px["object_id"] == MISS
    assert (got2[miss] == 0xFF123456).all() and (got2[~miss] == got[~miss]).all()
    om.destroy()


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_shade_all_matcaps(ctx, oracle, kind):
    w, h = 320, 200
    verts, tris, vc, v = _scene(20, w, h, j.DEFAULT_FLAGS, 12.0)
    om = oracle.mesh(verts, tris)
    px = oracle.cast([om], v)
    mc, cav = j.make_matcap(kind)
    bg = j.fill_background(w, h)
    want = oracle.shade(px, v, mc, cav, bg.copy())
    got = ctx.shade(px, v, mc, cav, background=bg)
    compare_rgba(got, want, tag=f"matcap {kind}")
    om.destroy()


@pytest.mark.parametrize("n,flags,scan_order", [
    (200003, j.DEFAULT_FLAGS, False), (50001, j.DEFAULT_FLAGS | j.ONE_BIT, False), (100002, j.DEFAULT_FLAGS & ~j.SHADING, False),
    (7, j.DEFAULT_FLAGS, False), (4, j.DEFAULT_FLAGS, False), (0, j.DEFAULT_FLAGS, False),
    (200003, j.DEFAULT_FLAGS, True), (100001, j.DEFAULT_FLAGS & ~j.SHADING, True),
])
def test_splat_matches_oracle(ctx, oracle, n, flags, scan_order):
    """Bit-exact splat, including the reference's order-dependent SIMD-packet collisions
    (render.h:783-807); scan_order sorts the cloud spatially so that nearly every packet collides."""
    w, h = 320, 180
    verts, tris = j.icosphere(8)
    verts = (verts * 0.6).astype(np.float32)
    pos, nrm, clr = j.cloud(max(n, 1))
    pos, nrm, clr = pos[:n].copy(), nrm[:n].copy(), clr[:n].copy()
    if scan_order:
        order = np.lexsort((pos[:, 0], pos[:, 1]))
        pos, nrm, clr = pos[order].copy(), nrm[order].copy(), clr[order].copy()
    mn, mx = j.compute_bb(np.concatenate([verts, pos]))
    v = j.orbit_view(j.make_view(w, h, mn, mx, flags), 15.0)
    om = oracle.mesh(verts, tris)
    px = oracle.cast([om], v)
    mc, cav = j.make_matcap(0)
    rgba0 = oracle.shade(px, v, mc, cav, j.fill_background(w, h))
    want_px, want_rgba = px.copy(), rgba0.copy()
    oracle.splat([(pos, nrm, clr, None, 0x40000000)], v, px, want_px, want_rgba)
    cl = ctx.cloud_create(pos, nrm, clr)
    got_px, got_rgba = px.copy(), rgba0.copy()
    ctx.splat([cl], v, px, got_px, got_rgba)
    is_pt = want_px["db_id"] == 0x40000000
    assert n == 0 or is_pt.sum() > 0
    assert (got_px["object_id"] == want_px["object_id"]).all()
    assert (got_px["db_id"] == want_px["db_id"]).all()
    assert (got_px["depth"] == want_px["depth"]).all()
    assert got_px.tobytes() == want_px.tobytes()
    assert (got_rgba == want_rgba).all()
    cl.destroy(); om.destroy()


def test_splat_two_clouds_and_transform(ctx, oracle):
    w, h = 256, 160
    p1, n1, c1 = j.cloud(30001, seed=1)
    p2, n2, c2 = j.cloud(20002, seed=2)
    cs = np.eye(4, dtype=np.float32)
    cs[3, :3] = [0.4, -0.2, 0.3]
    mn, mx = j.compute_bb(np.concatenate([p1, p2 + cs[3, :3]]))
    v = j.orbit_view(j.make_view(w, h, mn, mx, j.DEFAULT_FLAGS), 40.0)
    px = np.zeros((h, w), j.PIXEL_DTYPE)
    px["object_id"] = MISS
    px["depth"] = np.float32(FMAX)
    rgba0 = j.fill_background(w, h)
    want_px, want_rgba = px.copy(), rgba0.copy()
    oracle.splat([(p1, n1, c1, None, 0x40000000), (p2, n2, c2, cs.reshape(-1), 0x40000001)], v, px, want_px, want_rgba)
    a = ctx.cloud_create(p1, n1, c1, db_id=0x40000000)
    b = ctx.cloud_create(p2, n2, c2, cs=cs.reshape(-1), db_id=0x40000001)
    got_px, got_rgba = px.copy(), rgba0.copy()
    ctx.splat([a, b], v, px, got_px, got_rgba)
    assert got_px.tobytes() == want_px.tobytes()
    assert (got_rgba == want_rgba).all()
    assert ((got_px["db_id"] == 0x40000001).sum() > 100) and ((got_px["db_id"] == 0x40000000).sum() > 100)
    a.destroy(); b.destroy()


def test_render_frame_equals_stages(ctx, oracle):
    """j3dg_render_frame == cast -> shade -> splat composed by hand (view::render_scene order)."""
    w, h = 400, 240
    verts, tris, vc, v = _scene(20, w, h, j.DEFAULT_FLAGS | j.SHADOW, 25.0, True)
    pos, nrm, clr = j.cloud(60001)
    pos = (pos * 1.3).astype(np.float32)
    m = ctx.mesh_create(verts, tris, vcolors=vc)
    cl = ctx.cloud_create(pos, nrm, clr)
    mc, cav = j.make_matcap(0)
    px = ctx.cast([m], v)
    rgba = ctx.shade(px, v, mc, cav, background=j.fill_background(w, h))
    px2 = px.copy()
    ctx.splat([cl], v, px, px2, rgba)
    fpx = np.zeros((h, w), j.PIXEL_DTYPE)
    frgba = np.zeros((h, w), np.uint32)
    ctx.render_frame([m], [cl], v, mc, cav, pixels_out=fpx, rgba_out=frgba)
    assert fpx.tobytes() == px2.tobytes()
    assert (frgba == rgba).all()
    # and the whole frame against the oracle
    om = oracle.mesh(verts, tris, vcolors=vc)
    opx = oracle.cast([om], v)
    orgba = oracle.shade(opx, v, mc, cav, j.fill_background(w, h))
    opx2 = opx.copy()
    oracle.splat([(pos, nrm, clr, None, 0x40000000)], v, opx, opx2, orgba)
    compare_rgba(frgba, orgba, tag="frame")
    m.destroy(); cl.destroy(); om.destroy()


@pytest.mark.parametrize("w,h,world,flags", [
    (400, 240, 2, j.DEFAULT_FLAGS | j.SHADOW),
    (333, 201, 3, j.DEFAULT_FLAGS),                       # ragged: 6.3 bands over 3 ranks
    (640, 360, 8, j.DEFAULT_FLAGS | j.SHADOW),            # more ranks than some ranks have bands for
    (320, 200, 4, (j.DEFAULT_FLAGS | j.WIREFRAME) & ~j.EDGES),
])
def test_screen_shards_reassemble_to_the_full_frame(ctx, w, h, world, flags):
    """BASELINE configs[2] on one GPU: render the frame once unsharded and once per rank with
    j3dg_ctx_set_screen_shard(rank, world) into poisoned device buffers.  Every rank's own 32-row bands must be
    bit-identical to the unsharded frame (pixel records and RGBA — the halo row above each band feeds the edge
    shader), rows of other ranks must stay untouched except that halo row, and the bands of all ranks cover
    the frame exactly once (dist.bands_for_rank is what gather_bands ships)."""
    import torch
    from j3d_b200 import dist as jd
    verts, tris, vc, v = _scene(30, w, h, flags, 40.0, True)
    m = ctx.mesh_create(verts, tris, vcolors=vc)
    mc, cav = j.make_matcap(0)
    ctx.set_matcap(mc, cav)
    POISON = 0x5A
    def render():
        px = torch.full((h, w, 32), POISON, dtype=torch.uint8, device="cuda")
        rgba = torch.full((h, w, 4), POISON, dtype=torch.uint8, device="cuda")
        ctx.render_frame([m], [], v, pixels_out=px, rgba_out=rgba)
        ctx.synchronize()
        return px.cpu().numpy(), rgba.cpu().numpy()
    full_px, full_rgba = render()
    assert (full_px.view(j.PIXEL_DTYPE)["object_id"] != MISS).any()
    covered = np.zeros(h, np.int32)
    try:
        for rank in range(world):
            ctx.set_screen_shard(rank, world)
            px, rgba = render()
            own = np.zeros(h, bool)
            for (y0, y1) in jd.bands_for_rank(h, rank, world):
                own[y0:y1 + 1] = True
                covered[y0:y1 + 1] += 1
            assert px[own].tobytes() == full_px[own].tobytes()
            assert rgba[own].tobytes() == full_rgba[own].tobytes()
            halo = np.zeros(h, bool)
            halo[:-1] = own[1:] & ~own[:-1]
            untouched = ~own & ~halo
            assert (px[untouched] == POISON).all() and (rgba[~own] == POISON).all()
            tm = ctx.timings(reset=True)
            assert tm.kernel_launches > 0
    finally:
        ctx.set_screen_shard(0, 1)
    assert (covered == 1).all()
    with pytest.raises(j.J3dgError):
        ctx.set_screen_shard(2, 2)
    m.destroy()


def test_pipelined_frames_equal_synchronous(ctx):
    """j3dg_frame_submit / j3dg_frame_wait: every frame of a sweep equals the synchronous j3dg_render_frame."""
    verts, tris, vc, v0 = _scene(25, 320, 180, j.DEFAULT_FLAGS | j.SHADOW, 0.0, True)
    m = ctx.mesh_create(verts, tris, vcolors=vc)
    mc, cav = j.make_matcap(0)
    ctx.set_matcap(mc, cav)
    views = [j.orbit_view(v0, 17.0 * k) for k in range(5)]
    want = []
    for v in views:
        px = np.zeros((180, 320), j.PIXEL_DTYPE); rgba = np.zeros((180, 320), np.uint32)
        ctx.render_frame([m], [], v, pixels_out=px, rgba_out=rgba)
        want.append((px, rgba))
    got = [(np.zeros((180, 320), j.PIXEL_DTYPE), np.zeros((180, 320), np.uint32)) for _ in views]
    for k, v in enumerate(views):
        ctx.frame_submit([m], [], v, pixels_out=got[k][0], rgba_out=got[k][1])
        if k >= 1:
            ctx.frame_wait()
    ctx.frame_wait()
    for (gp, gr), (wp, wr) in zip(got, want):
        assert gp.tobytes() == wp.tobytes()
        assert (gr == wr).all()
    with pytest.raises(j.J3dgError):
        ctx.frame_wait()  # nothing in flight
    m.destroy()


def test_dirty_rect_readback_is_byte_identical(ctx):
    """j3dg_ctx_set_dirty_rect: persistent host buffers that receive only the changed rectangle end up byte-identical
    to full copies — object moving across the canvas, two alternating host buffers, pipelined and synchronous calls,
    a settings change, an empty scene and a background change in between."""
    w, h = 320, 180
    verts, tris = j.icosphere(12)
    verts = (verts * 0.35).astype(np.float32)
    mn, mx = np.array([-1.5, -1, -1], np.float32), np.array([1.5, 1, 1], np.float32)  # scene bbox larger than the object
    v0 = j.make_view(w, h, mn, mx, j.DEFAULT_FLAGS)
    mc, cav = j.make_matcap(0)
    ctx.set_matcap(mc, cav)
    m = ctx.mesh_create(verts, tris)

    def place(k):
        cs = np.eye(4, dtype=np.float32)
        cs[3, 0] = -1.2 + 0.45 * k  # column-major translation: the object walks from left to right
        cs[3, 1] = 0.3 * ((k % 3) - 1)
        m.set_cs(cs.reshape(-1))

    frames = []
    for k in range(7):
        v = v0.copy()
        if k == 4:
            v.flags = j.DEFAULT_FLAGS | j.WIREFRAME
        frames.append(v)

    def full(k, meshes, bg):
        place(k)
        px = np.zeros((h, w), j.PIXEL_DTYPE); rgba = np.zeros((h, w), np.uint32)
        ctx.render_frame(meshes, [], frames[k], pixels_out=px, rgba_out=rgba, bg_bottom=bg)
        return px, rgba

    ctx.set_dirty_rect(False)
    want = [full(k, [m] if k != 5 else [], 0xFF404040 if k != 6 else 0xFF804020) for k in range(7)]
    ctx.set_dirty_rect(True)
    try:
        bufs = [(np.zeros((h, w), j.PIXEL_DTYPE), np.zeros((h, w), np.uint32)) for _ in range(2)]
        # pipelined, buffers alternate
        for k in range(5):
            place(k)
            ctx.frame_submit([m], [], frames[k], pixels_out=bufs[k & 1][0], rgba_out=bufs[k & 1][1])
            if k >= 1:
                ctx.frame_wait()
                assert bufs[(k - 1) & 1][0].tobytes() == want[k - 1][0].tobytes(), k
                assert (bufs[(k - 1) & 1][1] == want[k - 1][1]).all(), k
        ctx.frame_wait()
        assert bufs[0][0].tobytes() == want[4][0].tobytes() and (bufs[0][1] == want[4][1]).all()
        # synchronous calls into the same buffers: empty scene (everything the buffer held must be erased), new background
        place(5)
        ctx.render_frame([], [], frames[5], pixels_out=bufs[1][0], rgba_out=bufs[1][1])
        assert bufs[1][0].tobytes() == want[5][0].tobytes() and (bufs[1][1] == want[5][1]).all()
        place(6)
        ctx.render_frame([m], [], frames[6], pixels_out=bufs[1][0], rgba_out=bufs[1][1], bg_bottom=0xFF804020)
        assert bufs[1][0].tobytes() == want[6][0].tobytes() and (bufs[1][1] == want[6][1]).all()
    finally:
        ctx.set_dirty_rect(False)
    m.destroy()


def test_degenerate_inputs(ctx, oracle):
    """Edge cases: single triangle, duplicated triangles, zero-area triangles, unreferenced vertices."""
    w, h = 160, 120
    v3 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, 5]], np.float32)
    for tris in (np.array([[0, 1, 2]], np.uint32),
                 np.array([[0, 1, 2]] * 9, np.uint32),
                 np.array([[0, 1, 2], [0, 0, 1], [1, 1, 1], [2, 1, 0]], np.uint32)):
        view = j.make_view(w, h, [0, 0, 0], [1, 1, 0.5])
        m = ctx.mesh_create(v3, tris)
        got = ctx.cast([m], view)
        om = oracle.mesh(v3, tris)
        want = oracle.cast([om], view)
        assert ((got["object_id"] != MISS) == (want["object_id"] != MISS)).all()
        hit = want["object_id"] != MISS
        assert hit.sum() > 0
        assert np.abs(got["depth"][hit] - want["depth"][hit]).max() <= 1e-5 * want["depth"][hit].max()
        m.destroy(); om.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["n2", "n256", "n257", "n513", "dup700", "cluster_outlier", "strip1025"])
def test_builder_block_boundaries_and_ties(ctx, oracle, case):
    """The builder works on blocks of 256 sorted triangles: sizes around the block boundaries, runs of identical Morton codes
    longer than a block (ties are broken by position; their nodes straddle blocks and go through the upper-tree kernels),
    a cluster with one far outlier (the sorted codes of one block differ in their top bits)."""
    rng = np.random.default_rng(7)
    w, h = 192, 128

    def soup(n, scale=1.0, centre=(0.0, 0.0, 0.0)):
        c = rng.uniform(-1, 1, (n, 1, 3)) * scale + np.asarray(centre)
        v = (c + rng.uniform(-0.08, 0.08, (n, 3, 3)) * scale).reshape(-1, 3).astype(np.float32)
        return v, np.arange(3 * n, dtype=np.uint32).reshape(n, 3)

    if case in ("n2", "n256", "n257", "n513"):
        v, t = soup(int(case[1:]))
    elif case == "dup700":
        v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0.2, 0.2, -0.5], [0.9, 0.1, -0.5], [0.1, 0.9, -0.5]], np.float32)
        t = np.array([[0, 1, 2]] * 700 + [[3, 4, 5]] * 330, np.uint32)
    elif case == "cluster_outlier":
        v0, t0 = soup(900, scale=0.15)
        v1, t1 = soup(3, scale=0.6, centre=(3.0, 3.0, 3.0))
        v = np.concatenate([v0, v1]); t = np.concatenate([t0, t1 + len(v0)])
    else:  # a long thin strip: a deep, unbalanced radix tree
        x = np.linspace(0, 50, 1026, dtype=np.float32)
        v = np.stack([np.stack([x, np.zeros_like(x), np.zeros_like(x)], 1), np.stack([x, np.ones_like(x), np.zeros_like(x)], 1)], 1).reshape(-1, 3)
        t = np.array([[2 * i, 2 * i + 2, 2 * i + 1] for i in range(1025)], np.uint32)
    mn, mx = v.min(0), v.max(0)
    view = j.make_view(w, h, mn, mx)
    m = ctx.mesh_create(v, t)
    assert m.info().nr_of_triangles == len(t)
    got = ctx.cast([m], view)
    om = oracle.mesh(v, t)
    want = oracle.cast([om], view)
    hit = want["object_id"] != MISS
    assert ((got["object_id"] != MISS) == hit).all()
    assert hit.sum() > 0
    assert np.abs(got["depth"][hit] - want["depth"][hit]).max() <= 1e-5 * np.abs(want["depth"][hit]).max()
    m.rebuild()
    again = ctx.cast([m], view)
    assert (again["depth"] == got["depth"]).all() and (again["object_id"] == got["object_id"]).all()
    m.destroy(); om.destroy()



def test_reference_library_agrees(ctx):
    """CUDA path vs the unmodified reference (prebuilt oracle/_ref), when it travelled along."""
    from oracle.bindings import Ref, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libj3d_ref.so not present on this box")
    w, h = 640, 360
    verts, tris = j.icosphere(59)  # config A mesh (69 620 triangles)
    ref = Ref(w, h)
    ref.add_mesh(verts, tris)
    ref.unzoom()
    v = ref.view()
    v.flags = j.DEFAULT_FLAGS | j.SHADOW
    ref.set_view(v)
    ref.render(7)
    want_px, want_rgba = ref.pixels(0), ref.image()
    m = ctx.mesh_create(verts, tris)
    mc, cav = j.make_matcap(0)
    got_px = np.zeros((h, w), j.PIXEL_DTYPE)
    got_rgba = np.zeros((h, w), np.uint32)
    ctx.render_frame([m], [], v, mc, cav, pixels_out=got_px, rgba_out=got_rgba)
    compare_pixels(got_px, want_px, tag="vs reference")
    compare_rgba(got_rgba, want_rgba, tag="vs reference")
    ref.close(); m.destroy()


# ---- BASELINE.json configs at their own sizes against the unmodified reference (oracle/_ref) --------------
def _ref_or_skip():
    from oracle.bindings import Ref, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libj3d_ref.so not present on this box")
    return Ref


@pytest.mark.timeout(600)
@pytest.mark.parametrize("flags", [j.DEFAULT_FLAGS, j.DEFAULT_FLAGS | j.SHADOW])
def test_config_a_1080p_matches_reference(ctx, flags):
    """BASELINE configs[0]: the 69 620-triangle mesh at 1920x1080, primary rays + matcap shading (and with shadow
    rays), pixel records and RGBA against the reference's own canvas (view.cpp:421-430)."""
    Ref = _ref_or_skip()
    w, h = 1920, 1080
    verts, tris = j.icosphere(59)
    ref = Ref(w, h)
    ref.add_mesh(verts, tris)
    ref.unzoom()
    v = ref.view()
    v.flags = flags
    ref.set_view(v)
    ref.render(3)
    want_px, want_rgba = ref.pixels(0), ref.image()
    m = ctx.mesh_create(verts, tris)
    mc, cav = j.make_matcap(0)
    got_px = np.zeros((h, w), j.PIXEL_DTYPE)
    got_rgba = np.zeros((h, w), np.uint32)
    ctx.render_frame([m], [], v, mc, cav, pixels_out=got_px, rgba_out=got_rgba)
    st = compare_pixels(got_px, want_px, tag="config A 1080p")
    assert st["hits"] > 500_000
    compare_rgba(got_rgba, want_rgba, tag="config A 1080p")
    ref.close(); m.destroy()


@pytest.mark.timeout(1800)
def test_config_b_1080p_matches_reference(ctx, config_b):
    """BASELINE configs[1], the headline: 28 037 120 triangles at 1920x1080 against the reference's pixel buffer and
    image, at the default pose and at an orbit pose with shadow rays — the deep-tree paths (tiny quantisation
    exponents, evictions at the lane budget, the 96-entry group stack) that only this size exercises."""
    Ref = _ref_or_skip()
    verts, tris, m, v0 = config_b
    w, h = 1920, 1080
    ref = Ref(w, h)
    ref.add_mesh(verts, tris)
    ref.unzoom()
    rv = ref.view()
    for name in ("near_plane", "diagonal"):
        assert getattr(rv, name) == getattr(v0, name)
    assert list(rv.cs) == list(v0.cs) and list(rv.projection_inv) == list(v0.projection_inv)
    mc, cav = j.make_matcap(0)
    for angle, flags in ((0.0, j.DEFAULT_FLAGS), (137.0, j.DEFAULT_FLAGS | j.SHADOW)):
        v = j.orbit_view(v0, angle) if angle else v0.copy()
        v.flags = flags
        ref.set_view(v)
        ref.render(3)
        want_px, want_rgba = ref.pixels(0), ref.image()
        got_px = np.zeros((h, w), j.PIXEL_DTYPE)
        got_rgba = np.zeros((h, w), np.uint32)
        ctx.render_frame([m], [], v, mc, cav, pixels_out=got_px, rgba_out=got_rgba)
        st = compare_pixels(got_px, want_px, tag=f"config B 1080p @{angle}")
        assert st["hits"] > 600_000
        compare_rgba(got_rgba, want_rgba, tag=f"config B 1080p @{angle}")
    ref.close()


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("n", [1_000_003, 10_000_001])
def test_splat_1080p_matches_reference(ctx, n):
    """BASELINE configs[3] scaled to what the reference's serial splat loop finishes in seconds: a vertex-coloured
    cloud of n points (n mod 4 != 0: the scalar tail path) over a mesh at 1920x1080.  The splat runs on the
    reference's own pre-splat canvas, so winners (point index, depth, db id) and colours must be bit-exact."""
    Ref = _ref_or_skip()
    w, h = 1920, 1080
    verts, tris = j.icosphere(59)
    verts = (verts * 0.8).astype(np.float32)
    pos, nrm, clr = j.cloud(n)
    ref = Ref(w, h)
    ref.add_mesh(verts, tris)
    ref.add_cloud(pos, nrm, clr)
    ref.unzoom()
    v = j.orbit_view(ref.view(), 25.0)
    ref.set_view(v)
    ref.render(3)
    px_in, rgba_in = ref.pixels(0), ref.image()
    ref.render(4)
    want_px, want_rgba = ref.pixels(1), ref.image()
    cl = ctx.cloud_create(pos, nrm, clr)
    got_px, got_rgba = px_in.copy(), rgba_in.copy()
    ctx.splat([cl], v, px_in, got_px, got_rgba)
    is_pt = want_px["db_id"] == 0x40000000
    assert is_pt.sum() > 100_000
    assert got_px.tobytes() == want_px.tobytes()
    assert (got_rgba == want_rgba).all()
    # and the whole frame through j3dg_render_frame: cast + shade + splat in one call
    m = ctx.mesh_create(verts, tris)
    mc, cav = j.make_matcap(0)
    f_px = np.zeros((h, w), j.PIXEL_DTYPE)
    f_rgba = np.zeros((h, w), np.uint32)
    ctx.render_frame([m], [cl], v, mc, cav, pixels_out=f_px, rgba_out=f_rgba)
    same_owner = (f_px["db_id"] == want_px["db_id"])
    assert same_owner.mean() > 0.9999
    pt = is_pt & same_owner
    assert (f_px["object_id"][pt] == want_px["object_id"][pt]).mean() > 0.9999
    compare_rgba(f_rgba, want_rgba, tag=f"frame with {n} points")
    ref.close(); cl.destroy(); m.destroy()


# ---- golden fixtures (produced by the unmodified reference; tests/golden/make_golden.py) --------
from golden_util import CASES, load  # noqa: E402


@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_golden(ctx, name):
    g = load(name)
    v = g["view"]
    m = ctx.mesh_create(g["verts"], g["tris"], vcolors=g["vc"])
    clouds = []
    if g["cloud"] is not None:
        clouds = [ctx.cloud_create(*g["cloud"])]
    mc, cav = j.make_matcap(0)
    px = ctx.cast([m], v)
    compare_pixels(px, g["pixels"], tag=name)
    if clouds:
        # splat in isolation on the reference's own pre-splat buffer: integer / index work, exact
        px_in = g["pixels"]
        rgba = ctx.shade(px_in, v, mc, cav, background=j.fill_background(v.width, v.height))
        after = px_in.copy()
        ctx.splat(clouds, v, px_in, after, rgba)
        want = g["pixels_after_splat"]
        for f in ("object_id", "db_id", "depth"):
            assert (after[f] == want[f]).all(), f
    else:
        rgba = ctx.shade(px, v, mc, cav, background=j.fill_background(v.width, v.height))
    compare_rgba(rgba, g["rgba"], tag=name)
    m.destroy()
    for c in clouds:
        c.destroy()


@pytest.mark.parametrize("name", CASES)
def test_pick_matches_golden(ctx, name):
    """j3dg_pick on the reference's own pixel buffer (uploaded to the device) answers every query — record,
    db id, world position, closest vertex, pivot — bit for bit like the reference's host functions."""
    import torch
    g = load(name)
    v = g["view"]
    m = ctx.mesh_create(g["verts"], g["tris"], vcolors=g["vc"])
    clouds = [ctx.cloud_create(*g["cloud"])] if g["cloud"] is not None else []
    canvas = g["pixels_after_splat"] if g["pixels_after_splat"] is not None else g["pixels"]
    d_canvas = torch.from_numpy(canvas.view(np.uint8).reshape(v.height, v.width, 32).copy()).cuda()
    got = ctx.pick([m], clouds, v, g["pick_xy"], pixels=d_canvas)
    assert got.tobytes() == g["picks"].tobytes()
    m.destroy()
    for c in clouds:
        c.destroy()


def test_pick_on_resident_canvas(ctx, oracle):
    """RGBA-only readback: render a frame without downloading the pixel records, then pick from the canvas that
    stayed in HBM; the answers equal the oracle's on the (separately downloaded) records."""
    w, h = 320, 200
    verts, tris = j.icosphere(16)
    pos, nrm, clr = j.cloud(20001)
    pos = (pos * 1.25).astype(np.float32)
    mn, mx = j.compute_bb(np.concatenate([verts, pos]))
    v = j.orbit_view(j.make_view(w, h, mn, mx, j.DEFAULT_FLAGS | j.SHADOW), 20.0)
    mc, cav = j.make_matcap(0)
    cs = np.eye(4, dtype=np.float32); cs[3, 0] = 0.05  # column-major: translation x
    m = ctx.mesh_create(verts, tris, cs=cs.reshape(-1))
    cl = ctx.cloud_create(pos, nrm, clr)
    rgba = np.zeros((h, w), np.uint32)
    ctx.render_frame([m], [cl], v, mc, cav, pixels_out=None, rgba_out=rgba)
    xy = np.stack(np.meshgrid(np.arange(-3, w + 4, 3), np.arange(-2, h + 3, 2)), -1).reshape(-1, 2).astype(np.int32)
    got = ctx.pick([m], [cl], v, xy)
    px = np.zeros((h, w), j.PIXEL_DTYPE)
    ctx.render_frame([m], [cl], v, mc, cav, pixels_out=px, rgba_out=rgba)
    om = oracle.mesh(verts, tris, cs=cs.reshape(-1))
    want = oracle.pick(px, v, [om], [(pos, None, 0x40000000)], xy)
    assert got.tobytes() == want.tobytes()
    assert (got["db_id"] == 0x20000000).sum() > 100 and (got["db_id"] == 0x40000000).sum() > 100
    # a canvas of another size is not silently reused
    v2 = j.make_view(64, 48, mn, mx)
    with pytest.raises(j.J3dgError):
        ctx.pick([m], [cl], v2, xy[:4])
    m.destroy(); cl.destroy(); om.destroy()


def _hit_sets(off, hits, ids):
    return [dict(zip(ids[off[k]:off[k + 1]].tolist(), hits[off[k]:off[k + 1]])) for k in range(len(off) - 1)]


def test_find_all_matches_golden(ctx):
    """j3dg_mesh_find_all against the reference's qbvh::find_all_triangles (golden): the same triangles per ray
    (>= 99.9 % of the rays; a grazing hit may fall on either side of the interval bounds), distance and
    barycentrics within 1e-5 on the common ones (exact division instead of rcpps + Newton-Raphson)."""
    import make_golden as mg
    verts, tris, rays = mg.allhits_inputs()
    z = np.load(mg.HERE / "allhits.npz")
    m = ctx.mesh_create(verts, tris)
    off, hits, ids = m.find_all(rays)
    got, want = _hit_sets(off, hits, ids), _hit_sets(z["offsets"], z["hits"], z["ids"])
    same = sum(1 for a, b in zip(got, want) if a.keys() == b.keys())
    assert same >= 0.999 * len(want), (same, len(want))
    assert abs(int(off[-1]) - int(z["offsets"][-1])) <= 3
    worst = 0.0
    for a, b in zip(got, want):
        for k in a.keys() & b.keys():
            worst = max(worst, float(np.abs(a[k][:3] - b[k][:3]).max() / max(1.0, abs(float(b[k][2])))))
    assert worst <= 1e-5, worst
    # closest of all hits in front of the origin == the closest-hit query
    fwd = rays[500:800].copy()
    ch, cid = m.find_closest(fwd)
    o2, h2, i2 = m.find_all(fwd)
    for k in range(fwd.shape[0]):
        seg = slice(o2[k], o2[k + 1])
        if ch[k, 3] == 1:
            j0 = np.argmin(h2[seg, 2])
            assert i2[seg][j0] == cid[k] or abs(h2[seg, 2][j0] - ch[k, 2]) <= 1e-6
        else:
            assert o2[k] == o2[k + 1]
    m.destroy()


def test_find_all_edge_cases(ctx):
    verts, tris = j.icosphere(4)
    m = ctx.mesh_create(verts, tris)
    off, hits, ids = m.find_all(np.zeros((0, 8), np.float32))
    assert off.tolist() == [0] and hits.shape[0] == 0
    miss = np.array([[5, 5, 5, 1, 0, 0, 0, FMAX]], np.float32)
    off, hits, ids = m.find_all(miss)
    assert off.tolist() == [0, 0]
    through = np.array([[-3, 0.01, 0.02, 1, 0, 0, 0, FMAX]] * 70, np.float32)  # more rays than one block
    off, hits, ids = m.find_all(through)
    assert (np.diff(off.astype(np.int64)) == 2).all() and (hits[:, 2] > 0).all()
    # capacity too small: the call fails loudly but reports the size needed
    import ctypes as C
    o = np.zeros((71,), np.uint32); h = np.zeros((4, 4), np.float32); i = np.zeros((4,), np.uint32); tot = C.c_uint32()
    rc = ctx._L.j3dg_mesh_find_all(m._h, through.ctypes.data_as(C.c_void_p), 70, o.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p),
                                   i.ctypes.data_as(C.c_void_p), 4, C.byref(tot))
    assert rc == -1 and tot.value == 140 and o[-1] == 140
    m.destroy()
    empty = ctx.mesh_create(verts, np.zeros((0, 3), np.uint32))
    off, hits, ids = empty.find_all(through)
    assert off[-1] == 0
    empty.destroy()


@pytest.mark.parametrize("name", ["vox_white", "vox_colors", "vox_texture"])
def test_voxelize_matches_golden(ctx, oracle, name):
    """j3dg_mesh_voxelize against the reference's .vox export (golden, produced by the unmodified write_vox):
    identical grid dimensions; occupancy and palette indices identical up to hit points that lie within rounding
    of a voxel face (exact division vs rcpps); where several colours fall into one voxel the largest index wins
    (= the oracle's vmax), which is one of the values the reference's racing threads can leave."""
    import make_golden as mg
    verts, tris, vc, uv, tex, max_dim = mg.voxel_inputs(name)
    want = np.load(mg.HERE / "voxels.npz")[name]
    m = ctx.mesh_create(verts, tris, vcolors=vc, uv=uv, texture=tex)
    got = m.voxelize(max_dim)
    assert got.shape == want.shape
    om = oracle.mesh(verts, tris, vcolors=vc, uv=uv, texture=tex)
    vmax, vmin = om.voxelize(max_dim)
    occupied = int((want != 0).sum())
    assert int(((got != 0) != (want != 0)).sum()) <= max(2, occupied // 2000)
    assert int((got != vmax).sum()) <= max(2, occupied // 2000)
    single = (vmin == vmax) & (got != 0) & (want != 0)
    assert int((got[single] != want[single]).sum()) <= max(2, occupied // 2000)
    m.destroy(); om.destroy()


def test_cpp_shim_renders_like_the_abi(ctx):
    """The C++ mirror of j3d's scene / canvas (j3dg_host.h), driven like view::load_*_from_file + view::render_scene by a
    g++-built program, produces byte-identical pixel records, image, picks and voxels to the ctypes path."""
    import subprocess
    from conftest import ROOT, build_cpp_shim_driver
    exe = build_cpp_shim_driver()
    w, h, max_dim = 333, 201, 40
    flags = j.DEFAULT_FLAGS | j.SHADOW
    verts, tris = j.icosphere(14)
    pos, nrm, clr = j.cloud(30002)
    pos = (pos * 1.2).astype(np.float32)
    inp, outp = ROOT / "build" / "tests" / "shim_in.bin", ROOT / "build" / "tests" / "shim_out.bin"
    with open(inp, "wb") as f:
        f.write(np.array([w, h, verts.shape[0], tris.shape[0], pos.shape[0], flags, max_dim], np.uint32).tobytes())
        for a in (verts, tris, pos, nrm, clr):
            f.write(np.ascontiguousarray(a).tobytes())
    res = subprocess.run([str(exe), str(inp), str(outp)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    raw = outp.read_bytes()
    o = 0
    got_px = np.frombuffer(raw, j.PIXEL_DTYPE, w * h, o).reshape(h, w); o += w * h * 32
    got_im = np.frombuffer(raw, np.uint32, w * h, o).reshape(h, w); o += w * h * 4
    got_picks = np.frombuffer(raw, j.PICK_DTYPE, 3, o); o += 3 * 64
    dim = np.frombuffer(raw, np.uint32, 3, o); o += 12
    got_vox = np.frombuffer(raw, np.uint8, int(dim.prod()), o).reshape(dim[2], dim[1], dim[0]); o += int(dim.prod())
    sweep_bad = int(np.frombuffer(raw, np.uint32, 1, o)[0])
    assert sweep_bad == 0, f"{sweep_bad} frames of the three-lane j3dg::sweep differ from the synchronous frames"
    # the same calls through ctypes: add_object + prepare_scene + unzoom, then cast -> shade -> splat on host buffers
    mn, mx = j.compute_bb(np.concatenate([verts, pos]))
    v = j.make_view(w, h, mn, mx, flags)
    mc, cav = j.make_matcap(0)
    m = ctx.mesh_create(verts, tris)
    cl = ctx.cloud_create(pos, nrm, clr)
    px = ctx.cast([m], v)
    pixels = px.copy()
    im = j.fill_background(w, h)
    ctx.shade(pixels, v, mc, cav, out=im)
    ctx.splat([cl], v, pixels, px, im)
    assert got_px.tobytes() == px.tobytes()
    assert (got_im == im).all()
    want_picks = ctx.pick([m], [cl], v, np.array([[w // 2, h // 2], [w // 3, h // 3], [-5, 2]], np.int32))
    assert got_picks.tobytes() == want_picks.tobytes()
    assert (got_picks["db_id"][:2] != 0).all() and got_picks["db_id"][2] == 0
    assert (got_vox == m.voxelize(max_dim)).all() and int((got_vox != 0).sum()) > 1000
    m.destroy(); cl.destroy()


def test_random_transformed_scenes_match_oracle(ctx, oracle):
    """Two meshes and a cloud with random rigid object matrices, random settings / poses / canvas sizes, whole frame in one
    call: pixel records within the parity bars, splat winners exact, image within 1 LSB, picks exact on the GPU's own
    records (the oracle side of this test is pinned on the reference by tests/test_oracle.py::
    test_oracle_equals_reference_random_scenes)."""
    from test_oracle import _rigid
    rng = np.random.default_rng(77)
    flag_sets = [j.DEFAULT_FLAGS, j.DEFAULT_FLAGS | j.SHADOW, j.EDGES | j.VERTEXCOLORS | j.SHADOW, j.DEFAULT_FLAGS | j.WIREFRAME,
                 j.DEFAULT_FLAGS | j.ONE_BIT, j.SHADING | j.VERTEXCOLORS]
    mc, cav = j.make_matcap(0)
    for trial in range(6):
        w, h = 4 * int(rng.integers(30, 90)), int(rng.integers(60, 200))
        va, ta = j.icosphere(int(rng.integers(4, 20)))
        vb, tb = j.icosphere(int(rng.integers(3, 10)))
        vb = (vb * 0.5).astype(np.float32)
        ca, cb, cc = _rigid(rng, 0.2), _rigid(rng, 0.8), _rigid(rng, 0.3)
        vca = j.vertex_colors(va) if trial % 2 == 0 else None
        pos, nrm, clr = j.cloud(int(rng.integers(2000, 20000)) | 1)
        pos = (pos * 1.1).astype(np.float32)
        ma = ctx.mesh_create(va, ta, vcolors=vca, cs=ca, db_id=0x20000000)
        mb = ctx.mesh_create(vb, tb, cs=cb, db_id=0x20000001)
        cl = ctx.cloud_create(pos, nrm, clr, cs=cc)
        oa = oracle.mesh(va, ta, vcolors=vca, cs=ca, db_id=0x20000000)
        ob = oracle.mesh(vb, tb, cs=cb, db_id=0x20000001)
        # scene bbox of the transformed objects, like prepare_scene
        def world(vs, cs):
            m = np.asarray(cs, np.float32).reshape(4, 4).T
            return vs @ m[:3, :3].T + m[:3, 3]
        allp = np.concatenate([world(va, ca), world(vb, cb), world(pos, cc)]).astype(np.float32)
        v = j.orbit_view(j.make_view(w, h, allp.min(0), allp.max(0), flag_sets[trial % len(flag_sets)]), float(rng.uniform(0, 360)))
        want = oracle.cast([oa, ob], v)
        want_rgba = oracle.shade(want, v, mc, cav, oracle.fill_background(w, h))
        want2 = want.copy()
        oracle.splat([(pos, nrm, clr, cc, 0x40000000)], v, want, want2, want_rgba)
        px = np.zeros((h, w), j.PIXEL_DTYPE)
        rgba = np.zeros((h, w), np.uint32)
        ctx.render_frame([ma, mb], [cl], v, mc, cav, pixels_out=px, rgba_out=rgba)
        mesh_px = (want2["db_id"] != 0x40000000) & (px["db_id"] != 0x40000000)
        compare_pixels(np.where(mesh_px, px, want2), want2, tag=f"random scene {trial}")
        pts = want2["db_id"] == 0x40000000
        assert pts.sum() > 20 and (px["db_id"][pts] == 0x40000000).mean() > 0.999
        both = pts & (px["db_id"] == 0x40000000)
        assert (px["object_id"][both] == want2["object_id"][both]).mean() > 0.999
        compare_rgba(rgba, want_rgba, tag=f"random scene {trial}")
        xy = np.stack([rng.integers(-3, w + 3, 200), rng.integers(-3, h + 3, 200)], 1).astype(np.int32)
        got = ctx.pick([ma, mb], [cl], v, xy)
        assert got.tobytes() == oracle.pick(px, v, [oa, ob], [(pos, cc, 0x40000000)], xy).tobytes()
        ma.destroy(); mb.destroy(); cl.destroy(); oa.destroy(); ob.destroy()


def test_peer_frames_protocol_single_gpu(ctx):
    """The NVLink peer-memory frame exchange (csrc/peer.cu, dist.PeerFrames) with one rank: the shade kernel renders
    into the exchange buffer, arrival / release flags order the steps, and the exchanged frame equals a direct render."""
    import torch
    import torch.distributed as dist
    from j3d_b200.dist import PeerFrames
    created = False
    if not dist.is_initialized():
        import socket
        sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
        created = True
    try:
        w, h = 256, 160
        verts, tris = j.icosphere(10)
        mn, mx = j.compute_bb(verts)
        v0 = j.make_view(w, h, mn, mx)
        mc, cav = j.make_matcap(0)
        ctx.set_matcap(mc, cav)
        m = ctx.mesh_create(verts, tris)
        pf = PeerFrames(ctx, h, w, torch.device("cuda", 0))
        px = torch.empty((h, w, 32), dtype=torch.uint8, device="cuda")
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
        snaps, wants = [], []
        try:
            # no host synchronisation inside the loop: slot k & 1 is rewritten by frame k + 2 while the consumer of
            # frame k (a device copy enqueued between arrive and release) may still be pending
            for step in range(7):
                v = j.orbit_view(v0, 10.0 * step)
                k = pf.begin()
                assert k == step
                ctx.render_frame([m], [], v, pixels_out=px, rgba_out=pf.target(k))
                pf.arrive(k)
                with torch.cuda.stream(stream):
                    snaps.append(pf.frames(k)[0].clone())
                pf.release(k)
            ctx.synchronize()
            for step in range(7):
                want = np.zeros((h, w), np.uint32)
                ctx.render_frame([m], [], j.orbit_view(v0, 10.0 * step), rgba_out=want)
                wants.append(want)
        finally:
            ctx.set_stream(0)
        for got, want in zip(snaps, wants):
            assert (got.cpu().numpy().view(np.uint32) == want).all()
        assert ctx.status() == 0
        assert not ctx.stream_wait_timed_out()
        pf.close()
        m.destroy()
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [1, 2])
def test_cpp_host_drives_the_group_abi(ctx, world):
    """Multi-GPU behind the C ABI (SURVEY §8b "Context / multi-GPU"): a g++-built host without Python or torch forks one
    process per GPU, creates the NCCL group, broadcasts the mesh rank 0 built, renders an orbit sweep into rank 0's
    frame buffer over peer memory (every frame consumed between arrival and release and compared with rank 0's own
    render of that pose) and one band-sharded frame with shadows.  world = 2 needs two GPUs (exit code 77 = skipped)."""
    import subprocess
    import torch
    from conftest import build_cpp_group_driver
    if world > torch.cuda.device_count():
        pytest.skip(f"needs {world} GPUs")
    exe = build_cpp_group_driver()
    res = subprocess.run([str(exe), str(world), "480", "270", "4"], capture_output=True, text=True, timeout=500)
    if res.returncode == 77:
        pytest.skip("fewer GPUs than ranks")
    assert res.returncode == 0, res.stdout + res.stderr
    assert "OK" in res.stdout


# ---- BASELINE.json full size (config B: 28 037 120 triangles, 1080p): size-independent properties ----
def _numpy_closest(verts, tris, org, d, t_near):
    """Brute-force restatement of the Woop test over ALL triangles for one ray (float32, unfused)."""
    f = np.float32
    org = org.astype(f); d = d.astype(f)
    a = np.abs(d)
    kz = 2
    if a[0] > a[1]:
        if a[0] > a[2]:
            kz = 0
    elif a[1] > a[2]:
        kz = 1
    kx = (kz + 1) % 3
    ky = (kx + 1) % 3
    if d[kz] < 0:
        kx, ky = ky, kx
    Sz = f(1.0) / d[kz]
    Sx = d[kx] * Sz
    Sy = d[ky] * Sz
    best_t, best_id = np.inf, -1
    CH = 4_000_000
    for s in range(0, tris.shape[0], CH):
        t = tris[s:s + CH]
        A = verts[t[:, 0]] - org
        B = verts[t[:, 1]] - org
        Cc = verts[t[:, 2]] - org
        Ax = A[:, kx] - Sx * A[:, kz]; Ay = A[:, ky] - Sy * A[:, kz]
        Bx = B[:, kx] - Sx * B[:, kz]; By = B[:, ky] - Sy * B[:, kz]
        Cx = Cc[:, kx] - Sx * Cc[:, kz]; Cy = Cc[:, ky] - Sy * Cc[:, kz]
        U = Cx * By - Cy * Bx; V = Ax * Cy - Ay * Cx; Wd = Bx * Ay - By * Ax
        inside = ((U <= 0) & (V <= 0) & (Wd <= 0)) | ((U >= 0) & (V >= 0) & (Wd >= 0))
        det = (U + V) + Wd
        ok = inside & (det != 0)
        idx = np.nonzero(ok)[0]
        if idx.size == 0:
            continue
        T = (U[idx] * (Sz * A[idx, kz]) + V[idx] * (Sz * B[idx, kz])) + Wd[idx] * (Sz * Cc[idx, kz])
        tt = T * (f(1.0) / det[idx])
        good = tt > t_near
        if good.any():
            k = np.argmin(np.where(good, tt, np.inf))
            if tt[k] < best_t:
                best_t, best_id = float(tt[k]), int(s + idx[k])
    return best_t, best_id


@pytest.fixture(scope="module")
def config_b(ctx):
    verts, tris = j.icosphere(1184)
    m = ctx.mesh_create(verts, tris)
    mn, mx = j.compute_bb(verts)
    v = j.make_view(1920, 1080, mn, mx)
    yield verts, tris, m, v
    m.destroy()


@pytest.mark.timeout(600)
def test_full_size_build_and_cast_properties(ctx, config_b):
    verts, tris, m, v = config_b
    info = m.info()
    assert info.nr_of_triangles == 28037120 and info.nr_of_nodes > 0
    assert info.build_ms < 100.0, f"BVH build took {info.build_ms} ms (target: sub-100 ms)"
    assert np.allclose(list(info.bbox_min), verts.min(0)) and np.allclose(list(info.bbox_max), verts.max(0))
    px = ctx.cast([m], v)
    hit = px["object_id"] != MISS
    assert 600_000 < hit.sum() < 900_000
    # barycentrics inside the triangle, ids in range, depth positive, normals unit-bounded
    assert (px["object_id"][hit] < tris.shape[0]).all()
    bu, bv = px["barycentric_u"][hit], px["barycentric_v"][hit]
    assert (bu >= -1e-5).all() and (bv >= -1e-5).all() and (bu + bv <= 1 + 1e-5).all()
    assert (px["depth"][hit] > v.diagonal / 100).all()
    assert (px["u"][hit] ** 2 + px["v"][hit] ** 2 <= 1 + 1e-5).all()
    # the hit point reconstructed from the barycentrics lies on the ray at parameter depth
    ys, xs = np.nonzero(hit)
    sel = np.random.default_rng(0).choice(len(ys), 2000, replace=False)
    ys, xs = ys[sel], xs[sel]
    t = tris[px["object_id"][ys, xs]]
    k = 1 - px["barycentric_u"][ys, xs] - px["barycentric_v"][ys, xs]
    p = verts[t[:, 0]] * k[:, None] + verts[t[:, 1]] * px["barycentric_u"][ys, xs][:, None] + verts[t[:, 2]] * px["barycentric_v"][ys, xs][:, None]
    pinv = np.array(list(v.projection_inv), np.float64).reshape(4, 4).T
    cs = np.array(list(v.cs), np.float64).reshape(4, 4).T
    sp = np.stack([2 * ((xs + 0.5) / 1920) - 1, 2 * ((ys + 0.5) / 1080) - 1, np.full(len(xs), v.near_plane), np.ones(len(xs))], 1)
    d = (pinv @ sp.T).T
    d[:, 3] = 0
    d = (cs @ d.T).T[:, :3]
    o = cs[:3, 3]
    q = o + d * px["depth"][ys, xs][:, None].astype(np.float64)
    assert np.abs(q - p).max() < 2e-5
    # idempotence: a second cast and a cast after a rebuild give the identical buffer
    px2 = ctx.cast([m], v)
    assert px2.tobytes() == px.tobytes()
    m.rebuild()
    px3 = ctx.cast([m], v)
    same = px3["object_id"] == px["object_id"]
    assert same.mean() > 0.9999 and (px3["depth"][same] == px["depth"][same]).all()


@pytest.mark.timeout(900)
def test_full_size_samples_against_brute_force(ctx, config_b):
    """A handful of pixels of the 28 M-triangle frame against an exhaustive float32 Woop test."""
    verts, tris, m, v = config_b
    px = ctx.cast([m], v)
    pinv = np.array(list(v.projection_inv), np.float32).reshape(4, 4).T
    cs = np.array(list(v.cs), np.float32).reshape(4, 4).T
    f = np.float32
    for (x, y) in [(960, 540), (700, 400), (1200, 800), (961, 131), (5, 5), (624, 540)]:
        sp = np.array([f(2) * ((f(x) + f(0.5)) / f(1920)) - f(1), f(2) * ((f(y) + f(0.5)) / f(1080)) - f(1), f(v.near_plane), f(1)], f)
        d = ((pinv[:, 0] * sp[0] + pinv[:, 1] * sp[1]) + pinv[:, 2] * sp[2]) + pinv[:, 3] * sp[3]
        d[3] = 0
        d = ((cs[:, 0] * d[0] + cs[:, 1] * d[1]) + cs[:, 2] * d[2]) + cs[:, 3] * d[3]
        t, tid = _numpy_closest(verts, tris, cs[:3, 3].copy(), d[:3], f(v.diagonal) / f(100))
        got = px[y, x]
        if tid < 0:
            assert got["object_id"] == MISS
        else:
            assert abs(got["depth"] - t) <= 1e-5 * t
            assert got["object_id"] == tid or abs(got["depth"] - t) <= 1e-6 * t


@pytest.mark.timeout(900)
def test_full_size_triangle_order_invariance(ctx, config_b):
    """The closest hit does not depend on the triangle order the builder sees: a shuffled copy of the
    mesh renders the same depth buffer, and its ids map back through the permutation."""
    verts, tris, m, v = config_b
    rng = np.random.default_rng(3)
    perm = rng.permutation(tris.shape[0])
    sh = np.ascontiguousarray(tris[perm])
    m2 = ctx.mesh_create(verts, sh)
    assert m2.info().build_ms < 150.0
    a = ctx.cast([m], v)
    b = ctx.cast([m2], v)
    hit = a["object_id"] != MISS
    assert ((b["object_id"] != MISS) == hit).mean() > 0.99999
    both = hit & (b["object_id"] != MISS)
    mapped = perm[b["object_id"][both]]
    same = mapped == a["object_id"][both]
    assert same.mean() > 0.9999
    assert np.abs(a["depth"][both] - b["depth"][both]).max() <= 1e-5 * a["depth"][both].max()
    m2.destroy()


@pytest.mark.timeout(900)
def test_full_size_all_hits_pick_and_voxel_properties(ctx, config_b):
    """The query clients on the 28 M-triangle mesh, through size-independent properties: a ray through a closed surface
    enters as often as it leaves; the nearest of all hits is the closest-hit answer; every voxel the export fills is
    touched by the surface and the filled shell is closed along the three ray directions; picks reproduce the records."""
    verts, tris, m, v = config_b
    rng = np.random.default_rng(3)
    n = 4096
    a = rng.normal(size=(n, 3)); a = 3.0 * a / np.linalg.norm(a, axis=1, keepdims=True)
    b = rng.normal(size=(n, 3)); b = 0.8 * b / np.linalg.norm(b, axis=1, keepdims=True) * rng.random((n, 1)) ** (1 / 3)
    rays = np.concatenate([a, b - a, np.zeros((n, 1)), np.full((n, 1), FMAX)], axis=1).astype(np.float32)
    off, hits, ids = m.find_all(rays)
    counts = np.diff(off.astype(np.int64))
    assert (counts >= 2).all() and (counts % 2 == 0).mean() > 0.999
    assert (ids < tris.shape[0]).all() and (hits[:, 2] > 0).all()
    ch, cid = m.find_closest(rays)
    assert (ch[:, 3] == 1).all()
    first = np.array([np.argmin(hits[off[k]:off[k + 1], 2]) + off[k] for k in range(n)])
    agree = ids[first] == cid
    assert agree.mean() > 0.999 and np.abs(hits[first, 2][agree] - ch[agree, 2]).max() <= 1e-6
    # every reported hit lies on its triangle: the barycentric point is on the ray at the reported distance
    sel = rng.choice(hits.shape[0], 3000, replace=False)
    ray_of = np.searchsorted(off, sel, side="right") - 1
    t = tris[ids[sel]]
    u, w = hits[sel, 0].astype(np.float64), hits[sel, 1].astype(np.float64)
    p = verts[t[:, 0]] * (1 - u - w)[:, None] + verts[t[:, 1]] * u[:, None] + verts[t[:, 2]] * w[:, None]
    q = rays[ray_of, :3].astype(np.float64) + rays[ray_of, 3:6].astype(np.float64) * hits[sel, 2][:, None].astype(np.float64)
    assert np.abs(p - q).max() < 2e-5
    # voxel export: a closed shell — along x every grid column that meets the solid sees filled voxels at both ends
    grid = m.voxelize(96)
    dz, dy, dx = grid.shape
    assert max(dx, dy, dz) == 96 and (grid[grid != 0] == 255).all()
    filled = grid != 0
    centre = np.zeros_like(filled); centre[dz // 4: 3 * dz // 4, dy // 4: 3 * dy // 4, :] = True
    cols = filled.any(axis=2)
    assert cols[dz // 4: 3 * dz // 4, dy // 4: 3 * dy // 4].all()             # every central column crosses the surface
    firsts = filled.argmax(axis=2); lasts = dx - 1 - filled[:, :, ::-1].argmax(axis=2)
    assert (lasts[cols] > firsts[cols]).mean() > 0.95                          # ... twice (front and back of the shell)
    assert 0.02 < filled.mean() < 0.2                                           # a shell, not a solid
    # picking on the resident canvas of a full-size frame
    mc, cav = j.make_matcap(0)
    px = np.zeros((1080, 1920), j.PIXEL_DTYPE)
    ctx.render_frame([m], [], v, mc, cav, pixels_out=px)
    xy = np.stack([rng.integers(0, 1920, 500), rng.integers(0, 1080, 500)], 1).astype(np.int32)
    got = ctx.pick([m], [], v, xy)
    assert got["pixel"].tobytes() == px[xy[:, 1], xy[:, 0]].tobytes()
    hit = got["db_id"] != 0
    assert hit.sum() > 100 and (got["closest_vertex"][hit] < verts.shape[0]).all()
    tri_of = tris[got["pixel"]["object_id"][hit]]
    assert (got["closest_vertex"][hit][:, None] == tri_of).any(axis=1).all()   # the closest vertex is a corner of the hit triangle
    assert np.isnan(got["world_pos"][~hit]).all() and np.isfinite(got["world_pos"][hit]).all()


def test_two_contexts_share_a_mesh_and_keep_two_frames_in_flight(ctx):
    """Frames in flight (include/j3dg.h, j3dg_render_frame): a second context of the same device renders the odd frames of
    a sweep on its own stream, from the SAME mesh handle, while the even frames run on the first — nothing is synchronised
    in between.  Every frame equals the frame rendered alone; the same through j3dg_frame_submit / j3dg_frame_wait."""
    import torch
    verts, tris = j.icosphere(150)
    mesh = ctx.mesh_create(verts, tris, vcolors=j.vertex_colors(verts))
    mn, mx = j.compute_bb(verts)
    w, h = 960, 540
    v0 = j.make_view(w, h, mn, mx, j.DEFAULT_FLAGS | j.SHADOW | j.VERTEXCOLORS)
    views = [j.orbit_view(v0, 11.0 * k) for k in range(12)]
    mc, cav = j.make_matcap(0)
    ctx.set_matcap(mc, cav)
    ctx2 = j.Context(0)
    ctx2.set_matcap(mc, cav)
    lanes = [ctx, ctx2]
    dev = torch.device("cuda", 0)
    want = []
    for v in views:
        px = np.zeros((h, w), j.PIXEL_DTYPE); rgba = np.zeros((h, w), np.uint32)
        ctx.render_frame([mesh], [], v, pixels_out=px, rgba_out=rgba)
        want.append((px, rgba))
    d_px = [torch.zeros((h, w, 32), dtype=torch.uint8, device=dev) for _ in views]
    d_rgba = [torch.zeros((h, w), dtype=torch.int32, device=dev) for _ in views]
    for rep in range(3):
        for k, v in enumerate(views):  # device outputs: the calls return before the kernels ran
            lanes[k & 1].render_frame([mesh], [], v, pixels_out=d_px[k], rgba_out=d_rgba[k])
    for c in lanes:
        c.synchronize()
    for k in range(len(views)):
        assert d_px[k].cpu().numpy().tobytes() == want[k][0].tobytes(), k
        assert d_rgba[k].cpu().numpy().view(np.uint32).tobytes() == want[k][1].tobytes(), k
    # pipelined host frames on both contexts
    hpx = [[np.zeros((h, w), j.PIXEL_DTYPE) for _ in range(2)] for _ in range(2)]
    hrgba = [[np.zeros((h, w), np.uint32) for _ in range(2)] for _ in range(2)]
    done = []
    for k, v in enumerate(views):
        ln, b = k & 1, (k >> 1) & 1
        lanes[ln].frame_submit([mesh], [], v, pixels_out=hpx[ln][b], rgba_out=hrgba[ln][b])
        if k >= 2:
            lanes[ln].frame_wait()
            kk = k - 2
            done.append((kk, hpx[kk & 1][(kk >> 1) & 1].copy(), hrgba[kk & 1][(kk >> 1) & 1].copy()))
    for kk in (len(views) - 2, len(views) - 1):
        lanes[kk & 1].frame_wait()
        done.append((kk, hpx[kk & 1][(kk >> 1) & 1].copy(), hrgba[kk & 1][(kk >> 1) & 1].copy()))
    assert sorted(d[0] for d in done) == list(range(len(views)))
    for kk, px, rgba in done:
        assert px.tobytes() == want[kk][0].tobytes() and rgba.tobytes() == want[kk][1].tobytes(), kk
    assert ctx.status() == 0 and ctx2.status() == 0
    ctx2.close()
    mesh.destroy()


@pytest.mark.timeout(900)
def test_many_objects_through_the_top_level_tree(ctx, oracle):
    """a8 (jtk/qbvh.h:3251-3387): scenes with many objects are cast through a per-frame top-level tree over the objects'
    world boxes instead of a loop over all of them.  1 200 rigidly placed small meshes (some overlapping), primary and
    shadow rays: the frame equals the oracle's (which loops over the objects) within the parity bars, and equals the
    frame of the linear loop of the same kernels (a second context with the tree switched off) except at exact ties
    between objects."""
    import os
    from test_oracle import _rigid
    rng = np.random.default_rng(123)
    w, h = 480, 272
    base_v, base_t = j.icosphere(3)
    n_obj = 1200
    gm, om, pts = [], [], []
    for k in range(n_obj):
        cs = np.asarray(_rigid(rng, 1.0), np.float32).copy()
        cs[12:15] = rng.uniform(-6.0, 6.0, 3).astype(np.float32)   # column-major translation
        scale = np.float32(rng.uniform(0.15, 0.6))
        vk = (base_v * scale).astype(np.float32)
        gm.append(ctx.mesh_create(vk, base_t, cs=cs, db_id=0x20000000 + k))
        om.append(oracle.mesh(vk, base_t, cs=cs, db_id=0x20000000 + k))
        m = cs.reshape(4, 4).T
        pts.append(vk @ m[:3, :3].T + m[:3, 3])
    allp = np.concatenate(pts).astype(np.float32)
    mc, cav = j.make_matcap(0)
    old = os.environ.get("J3DG_TOP_MIN")
    os.environ["J3DG_TOP_MIN"] = "1000000"
    linear = j.Context(0)   # the same kernels, objects looped
    if old is None:
        del os.environ["J3DG_TOP_MIN"]
    else:
        os.environ["J3DG_TOP_MIN"] = old
    for angle, flags in ((20.0, j.DEFAULT_FLAGS | j.SHADOW), (200.0, j.DEFAULT_FLAGS)):
        v = j.orbit_view(j.make_view(w, h, allp.min(0), allp.max(0), flags), angle)
        want = oracle.cast(om, v)
        want_rgba = oracle.shade(want, v, mc, cav, oracle.fill_background(w, h))
        px = np.zeros((h, w), j.PIXEL_DTYPE); rgba = np.zeros((h, w), np.uint32)
        ctx.render_frame(gm, [], v, mc, cav, pixels_out=px, rgba_out=rgba)
        st = compare_pixels(px, want, tag=f"1200 objects @{angle}")
        assert st["hits"] > 0.1 * w * h
        compare_rgba(rgba, want_rgba, tag=f"1200 objects @{angle}")
        px2 = np.zeros((h, w), j.PIXEL_DTYPE); rgba2 = np.zeros((h, w), np.uint32)
        linear.render_frame(gm, [], v, mc, cav, pixels_out=px2, rgba_out=rgba2)
        same = (px["db_id"] == px2["db_id"]) & (px["object_id"] == px2["object_id"])
        assert same.mean() > 0.9995, same.mean()
        assert (px["depth"][same] == px2["depth"][same]).all() and (px["mark"][same] == px2["mark"][same]).mean() > 0.9995
        t_tree, t_lin = ctx.timings(reset=True).cast_ms, linear.timings(reset=True).cast_ms
        print(f"1200 objects @{angle}: cast {t_tree:.3f} ms through the tree, {t_lin:.3f} ms looping")
    linear.close()
    for m in gm:
        m.destroy()
    for m in om:
        m.destroy()
