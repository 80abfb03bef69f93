"""CPU tests of the procedural inputs."""
import numpy as np

import j3d_b200 as j


def test_icosphere_counts_and_topology(native):
    for f in (1, 2, 5, 17):
        verts, tris = j.icosphere(f)
        assert tris.shape[0] == 20 * f * f and verts.shape[0] == 10 * f * f + 2
        assert len(np.unique(tris)) == verts.shape[0]
        # closed 2-manifold: every undirected edge is shared by exactly two triangles, with opposite orientation
        e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]]).astype(np.int64)
        key = e[:, 0] * (1 << 32) + e[:, 1]
        rev = e[:, 1] * (1 << 32) + e[:, 0]
        assert len(np.unique(key)) == len(key)
        assert np.array_equal(np.sort(key), np.sort(rev))
        v0, v1, v2 = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
        n = np.cross(v1 - v0, v2 - v0)
        assert (np.einsum("ij,ij->i", n, (v0 + v1 + v2) / 3) > 0).all()
        r = np.linalg.norm(verts, axis=1)
        assert r.min() > 0.9 and r.max() < 1.1


def test_config_sizes(native):
    import ctypes as C
    S = j.capi.synth()
    for f, nt in ((59, 69620), (1184, 28037120), (3873, 300002580)):
        a, b = C.c_uint64(), C.c_uint64()
        S.synth_icosphere_counts(f, C.byref(a), C.byref(b))
        assert b.value == nt and a.value == nt // 2 + 2


def test_cloud_is_deterministic_and_range_independent(native):
    p, n, c = j.cloud(1000)
    p2, n2, c2 = j.cloud(400, first=300)
    assert np.array_equal(p[300:700], p2) and np.array_equal(n[300:700], n2) and np.array_equal(c[300:700], c2)
    assert (c >> 24 == 0xFF).all()
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)


def test_shuffle_is_a_permutation(native):
    verts, tris = j.icosphere(6)
    _, sh = j.icosphere(6, shuffle_seed=99)
    assert not np.array_equal(tris, sh)
    assert np.array_equal(np.sort(tris.view([("a", "u4"), ("b", "u4"), ("c", "u4")]).ravel()), np.sort(sh.view([("a", "u4"), ("b", "u4"), ("c", "u4")]).ravel()))
