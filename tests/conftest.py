import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def native():
    """Build (if needed) every native artefact that can be built on this machine."""
    from j3d_b200 import build
    build.build_synth()
    build.build_host()
    build.build_oracle()
    if not (ROOT / "j3d_b200" / "libj3dg.so").exists():
        build.build_cuda()
    return True


@pytest.fixture(scope="session")
def oracle(native):
    from oracle.bindings import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ctx(native):
    import j3d_b200 as j
    c = j.Context(0)  # raises loudly without a B200: there is no fallback path
    yield c
    c.close()


def build_cpp_shim_driver():
    """g++ build of tests/cpp/shim_frame.cpp (the C++ mirror of j3d's scene / canvas driven like view::render_scene)
    against libj3dg.so + libj3dg_host.so.  Returns the path of the executable."""
    import subprocess
    out = ROOT / "build" / "tests"
    out.mkdir(parents=True, exist_ok=True)
    exe = out / "shim_frame"
    pkg = ROOT / "j3d_b200"
    cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-I", str(pkg / "host"),
           str(ROOT / "tests" / "cpp" / "shim_frame.cpp"), "-o", str(exe), "-L", str(pkg), "-lj3dg", "-lj3dg_host", f"-Wl,-rpath,{pkg}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def build_cpp_group_driver():
    """g++ build of tests/cpp/group_ranks.cpp: a plain C++ host (no Python, no torch) that forks one process per GPU and
    drives the multi-GPU part of the C ABI (j3dg_group_*, j3dg_frames_*).  Returns the path of the executable."""
    import subprocess
    out = ROOT / "build" / "tests"
    out.mkdir(parents=True, exist_ok=True)
    exe = out / "group_ranks"
    pkg = ROOT / "j3d_b200"
    cuda = Path("/usr/local/cuda")
    cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-I", str(pkg / "host"),
           "-I", str(cuda / "include"), str(ROOT / "tests" / "cpp" / "group_ranks.cpp"), "-o", str(exe), "-L", str(pkg), "-lj3dg", "-lj3dg_host",
           "-L", str(cuda / "lib64"), "-lcudart", f"-Wl,-rpath,{pkg}", f"-Wl,-rpath,{cuda / 'lib64'}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe
