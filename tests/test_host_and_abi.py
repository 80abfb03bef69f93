"""CPU tests of the host logic and of the C-ABI library itself (no compute without a GPU)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import j3d_b200 as j
from j3d_b200 import capi

ROOT = Path(__file__).resolve().parents[1]


def test_abi_exports_every_declared_symbol(native):
    hdr = (ROOT / "include" / "j3dg.h").read_text()
    declared = sorted(set(re.findall(r"\b(j3dg_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 20
    L = capi.lib()  # loads libj3dg.so (static cudart) without needing a GPU
    for name in declared:
        assert hasattr(L, name), f"libj3dg.so does not export {name}"
    nm = subprocess.run(["nm", "-D", "--defined-only", str(ROOT / "j3d_b200" / "libj3dg.so")], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (j3dg_[a-z_0-9]+)", nm))
    assert set(declared) <= exported
    # nothing but the C ABI leaks out of the library
    assert all(s.startswith("j3dg_") for s in re.findall(r"\bT (\S+)", nm)), "non-ABI symbols are exported"


def test_struct_layouts():
    assert C.sizeof(j.View) == 4 * 4 + 4 * 64 + 12 + 4
    assert j.PIXEL_DTYPE.itemsize == 32
    assert [j.PIXEL_DTYPE.fields[k][1] for k in ("mark", "u", "v", "depth", "object_id", "barycentric_u", "barycentric_v", "db_id")] == [0, 4, 8, 12, 16, 20, 24, 28]


def test_no_gpu_fails_loudly(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(j.J3dgError) as e:
        j.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_oracle():
    """The product path must never import, link or load anything under oracle/."""
    for p in list((ROOT / "j3d_b200").rglob("*.py")) + list((ROOT / "j3d_b200").rglob("*.cu")) + list((ROOT / "j3d_b200").rglob("*.cuh")) + \
            list((ROOT / "j3d_b200").rglob("*.cpp")) + list((ROOT / "j3d_b200").rglob("*.h")):
        if p.name == "build.py":
            continue  # builds the checker, does not use it
        txt = p.read_text()
        assert "oracle" not in txt.lower() or "no oracle" in txt.lower() or p.name == "capi.py" and "touches oracle" in txt, p
    so = ROOT / "j3d_b200" / "libj3dg.so"
    if so.exists():
        ldd = subprocess.run(["ldd", str(so)], capture_output=True, text=True).stdout
        assert "oracle" not in ldd and "j3d_ref" not in ldd


def test_host_orbit_is_rigid_and_periodic(native):
    verts, _ = j.icosphere(4)
    mn, mx = j.compute_bb(verts)
    v0 = j.make_view(640, 360, mn, mx)
    for ang in (0.0, 17.0, 90.0, 251.0):
        v = j.orbit_view(v0, ang)
        cs = np.array(list(v.cs), np.float32).reshape(4, 4).T
        ci = np.array(list(v.cs_inv), np.float32).reshape(4, 4).T
        assert np.allclose(cs @ ci, np.eye(4), atol=1e-5)
        r = cs[:3, :3]
        assert np.allclose(r @ r.T, np.eye(3), atol=1e-5)
        # the pivot keeps its camera-space position (orbit about the pivot)
        p = np.array(list(v0.pivot) + [1.0], np.float32)
        assert np.allclose(ci @ p, np.array(list(v0.cs_inv), np.float32).reshape(4, 4).T @ p, atol=1e-4)
    a, b = j.orbit_view(v0, 0.0), v0
    assert np.allclose(list(a.cs), list(b.cs), atol=1e-6)


def test_background_and_matcap_shapes(native):
    bg = j.fill_background(64, 36)
    assert bg[0, 0] == 0xFF000000 and (bg >> 24 == 0xFF).all() and bg[-1, 0] != bg[0, 0]
    mc, cav = j.make_matcap(0)
    assert mc.shape == (512, 512) and cav == 0xFF7D7DFF


def test_cpp_shim_compiles_and_links(native):
    """The header-only C++ mirror of scene / canvas (j3d_b200/host/j3dg_host.h) builds warning-free against the C ABI and
    fails loudly without a GPU (no CPU path)."""
    import torch
    from conftest import build_cpp_shim_driver
    exe = build_cpp_shim_driver()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: tests/test_gpu_parity.py runs the driver for real")
    inp = ROOT / "build" / "tests" / "empty.bin"
    inp.write_bytes(np.array([16, 16, 0, 0, 0, 0, 4], np.uint32).tobytes())
    res = subprocess.run([str(exe), str(inp), str(inp.with_suffix(".out"))], capture_output=True, text=True)
    assert res.returncode == 1 and "j3dg_ctx_create" in res.stderr
