"""In-tree build of every native artefact (no JIT cache: the built .so files travel with
the repo snapshot to the GPU box).

  libj3dg.so        j3d_b200/csrc/*.cu   nvcc, sm_100a only      (the product)
  libj3dg_host.so   j3d_b200/host/*.cpp  g++                      (host-side mirror of j3d's canvas/scene/camera/matcap)
  libj3d_synth.so   j3d_b200/synth/synth.c  gcc -fopenmp          (procedural inputs)
  oracle/libj3d_oracle.so, oracle/_ref/libj3d_ref.so  via oracle/Makefile (test infrastructure)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
HOST = PKG / "host"
SYNTH = PKG / "synth"
ORACLE = ROOT / "oracle"
REF_ROOT = Path(os.environ.get("J3D_REF", "/root/reference"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def _run(cmd, log: Path | None = None):
    res = subprocess.run([str(c) for c in cmd], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        log.write_text(res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("build command failed: " + " ".join(str(c) for c in cmd))
    return res.stdout


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    out = PKG / "libj3dg.so"
    srcs = sorted(CSRC.glob("*.cu"))
    deps = srcs + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "j3dg.h"]
    if force or _newer(out, deps):
        cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-I", ROOT / "include", "-I", CSRC, "-o", out, *srcs]
        txt = _run(cmd, PKG / "build_ptxas.log")
        if verbose:
            print(txt)
    return out


def build_host(force: bool = False) -> Path:
    out = PKG / "libj3dg_host.so"
    srcs = sorted(HOST.glob("*.cpp"))
    deps = srcs + sorted(HOST.glob("*.h")) + [ROOT / "include" / "j3dg.h"]
    if force or _newer(out, deps):
        # -ffp-contract=off: the camera / pose arithmetic must round like the reference's
        # (plain SSE, no FMA) so views are bit-identical.
        _run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
              "-I", ROOT / "include", "-I", HOST, "-o", out, *srcs, "-ldl"])
    return out


def build_synth(force: bool = False) -> Path:
    out = PKG / "libj3d_synth.so"
    src = SYNTH / "synth.c"
    if force or _newer(out, [src]):
        _run(["gcc", "-std=c11", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-Wall",
              "-o", out, src, "-lm"])
    return out


def build_oracle(force: bool = False):
    """The checker: plain-C restatement always; the real reference when its sources exist."""
    if force:
        _run(["make", "-C", ORACLE, "clean"])
    _run(["make", "-C", ORACLE, "oracle"])
    if (REF_ROOT / "j3d" / "canvas.cpp").exists():
        _run(["make", "-C", ORACLE, "-j8", "ref", f"J3D_REF={REF_ROOT}"])
    return ORACLE / "libj3d_oracle.so"


def build_all(force: bool = False, verbose: bool = False):
    build_synth(force)
    build_host(force)
    build_oracle(force)
    build_cuda(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", *(str(p) for p in [PKG / "libj3dg.so", PKG / "libj3dg_host.so", PKG / "libj3d_synth.so"]))
