"""Loads the committed golden fixtures (tests/golden, produced from the unmodified reference by
tests/golden/make_golden.py) and regenerates their procedural inputs."""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "golden"))
import make_golden as mg  # noqa: E402
import j3d_b200 as j  # noqa: E402

META = json.loads((HERE / "golden" / "golden.json").read_text())
CASES = [k for k in mg.cases().keys()]


def view_from_dict(d) -> j.View:
    v = j.View()
    for k, val in d.items():
        if isinstance(val, list):
            arr = getattr(v, k)
            for i, x in enumerate(val):
                arr[i] = x
        else:
            setattr(v, k, val)
    return v


def load(name):
    """-> dict(case, view, verts, tris, vc, cloud, pixels, rgba, pixels_after_splat|None)"""
    meta = META[name]
    c = meta["case"]
    verts, tris, vc, cl = mg.inputs(c)
    got_crc = mg.crc(verts, tris, *([vc] if vc is not None else []), *(cl if cl else []))
    assert got_crc == meta["input_crc"], f"procedural input of golden case {name} drifted"
    z = np.load(HERE / "golden" / f"{name}.npz")
    px = np.ascontiguousarray(z["pixels"]).view(j.PIXEL_DTYPE).reshape(c["h"], c["w"])
    after = None
    if "pixels_after_splat" in z.files:
        after = np.ascontiguousarray(z["pixels_after_splat"]).view(j.PIXEL_DTYPE).reshape(c["h"], c["w"])
    pz = np.load(HERE / "golden" / "picks.npz")
    picks = np.ascontiguousarray(pz[name]).view(j.PICK_DTYPE).reshape(-1)
    return dict(case=c, view=view_from_dict(meta["view"]), verts=verts, tris=tris, vc=vc, cloud=cl,
                pixels=px, rgba=np.ascontiguousarray(z["rgba"]), pixels_after_splat=after,
                pick_xy=mg.pick_queries(c["w"], c["h"]), picks=picks)
