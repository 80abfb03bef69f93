#!/usr/bin/env python
"""bench.py — headline benchmark of the j3d hot path on B200 (contract: see DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload B|A|C|P]

Workload B (default; BASELINE.json configs[1] + configs[4]): synthetic 28 037 120-triangle noised geodesic
icosphere ("Lucy scale", f = 1184), 1920x1080, default settings (edges + matcap shading), default camera +
unzoom pose.  A *step* is one frame of the hot path per rank: ray cast (one primary ray per pixel, misses
included) + fused shading, the camera orbiting so successive frames touch different parts of the 1.6 GB BVH
(inputs larger than the 126 MB L2; nothing is cached between steps).  metric = primary Mrays/s.

N > 1 (torchrun, one rank per GPU): the mesh is replicated, the BVH is built once on rank 0 and broadcast over
NCCL, frames are sharded round-robin, every frame's RGBA ends up in rank 0's HBM (stores over NVLink peer memory).
  * `value` (weak scaling, `steps` frames per rank): the N ranks render an N-times FINER sweep of the SAME arc
    (frame i at i / N degrees, rank i mod N), so every N sees the same poses' cost — round 1 strode the orbit by N
    degrees and measured view cost instead of communication.
  * `sweep360` (strong scaling, BASELINE configs[4]): the fixed 360-frame orbit, frame k on rank k mod N, identical
    poses for every N; device-resident and end-to-end figures.
  * `e2e`: the same frames through the pipelined C-ABI calls with pinned HOST buffers, in the interactive-host mode
    the device-side picking enables (RGBA crosses PCIe, pixel records stay in HBM, j3dg_pick on demand);
    `e2e_full_records` is the mode that also downloads the 32-byte records every frame.

Workload A: the same on the 69 620-triangle mesh (configs[0]).  Workload C (configs[2]): ONE 3840x2160 frame with
shadow rays of the 300 M-triangle mesh, screen bands sharded over the N ranks into ONE frame in rank 0's HBM.
Workload P (configs[3]): 100 M-point depth splat at 1080p.

--impl reference: the reference's own std::thread CPU renderer (oracle/_ref, compiled from the unmodified j3d
sources) on the same workload, on this box's host cores, bounded to a few frames.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

SHADOW_FLAG = 1 << 1
WORKLOADS = {
    "A": dict(f=59, w=1920, h=1080, shadow=False),
    "B": dict(f=1184, w=1920, h=1080, shadow=False),
    "C": dict(f=3873, w=3840, h=2160, shadow=True),
    "P": dict(points=100_000_000, w=1920, h=1080),
}
SPLAT_BYTES_PER_POINT = 12  # the position read; normals / colours only for the <= W*H winners


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = Path(f"/tmp/j3dg_clocks_{os.getpid()}.csv")

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.path.read_text().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nme, val in zip(names, c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        try:
            self.path.unlink()
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def pin_to_gpu_numa(gpu_index: int):
    """Run this process (and allocate its pinned host buffers) on the CPUs next to its GPU: with 8 ranks the
    device->host streams otherwise land on whichever NUMA node the scheduler picked.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def workload_name(wl, f, nt, w, h):
    if wl == "C":
        return f"noised geodesic icosphere f={f} ({nt} triangles), {w}x{h}, 1 primary ray/pixel + 1 shadow ray/hit pixel + fused shading, one frame band-sharded over the ranks"
    return f"noised geodesic icosphere f={f} ({nt} triangles), {w}x{h}, 1 primary ray/pixel + fused shading, orbit 1 deg/step"


# ----------------------------------------------------------------------------------------------
# the reference on the host cores (cpu_baseline leg and --impl reference)
# ----------------------------------------------------------------------------------------------
def reference_frames(j, verts, tris, w, h, views, warm=1, keep_last=False):
    """add_object once, then the given frames (cast + canvas_to_image).  Returns a dict of timings (+ the last frame)."""
    from oracle.bindings import Ref
    ref = Ref(w, h)
    ref.add_mesh(verts, tris)
    build_s = ref.times()["build"]
    ref.unzoom()
    cast, shade = [], []
    for k, v in enumerate(views):
        ref.set_view(v)
        ref.render(3)
        t = ref.times()
        if k >= warm:
            cast.append(t["cast"]); shade.append(t["shade"])
    out = {"cores": ref.cores(), "build_s": build_s, "cast_s": cast, "shade_s": shade}
    if keep_last:
        out["pixels"], out["image"] = ref.pixels(0), ref.image()
    ref.close()
    return out


def shadow_rays_of(px):
    return int((px["object_id"] != 0xFFFFFFFF).sum())


def run_reference(args, wl, rank, world):
    if rank != 0:
        return 0
    import j3d_b200 as j
    from oracle.bindings import Ref, ref_available
    if not ref_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libj3d_ref.so was not built (reference sources absent at build time)"}))
        return 0
    cfg = WORKLOADS[wl]
    w, h = cfg["w"], cfg["h"]
    if wl == "P":
        n = min(args.points or cfg["points"], 10_000_000)  # bounded sample: the reference's splat loop is serial
        pos, nrm, clr = j.cloud(n)
        ref = Ref(w, h)
        ref.add_cloud(pos, nrm, clr)
        ref.unzoom()
        v0 = ref.view()
        ts = []
        for k in range(1 + min(args.steps, 5)):
            ref.set_view(j.orbit_view(v0, float(k)))
            ref.render(7)
            if k >= 1:
                ts.append(ref.times()["splat"])
        value = n * len(ts) / sum(ts) / 1e6
        sample = f"{n}-point sample of the vertex-coloured cloud, {len(ts)} frames {w}x{h} (render_pointclouds_on_image), 1 warm-up frame"
        line = {"impl": "reference", "metric": "Mpoints/s splatted @1080p", "value": value, "unit": "Mpoints/s", "n_gpus": args.gpus, "steps": len(ts),
                "warmup": 1, "ms_per_step": 1e3 * sum(ts) / len(ts), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": f"{args.points or cfg['points']}-point vertex-coloured cloud (normals + colours), {w}x{h}, empty mesh scene, orbit 1 deg/frame", "l2": "inputs_larger_than_l2"},
                "cpu_baseline": {"value": value, "unit": "Mpoints/s", "cores": ref.cores(), "kind": "reference", "sample": sample},
                "e2e": {"value": value, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0
    f = args.f or cfg["f"]
    verts, tris = j.icosphere(f)
    mn, mx = j.compute_bb(verts)
    flags = j.DEFAULT_FLAGS | (j.SHADOW if cfg["shadow"] else 0)
    v0 = j.make_view(w, h, mn, mx, flags)
    steps = min(args.steps, 3 if wl == "C" else 12)  # bounded sample: a CPU frame of the 28M mesh takes ~0.1-1 s
    step_deg = 25.0 if wl == "C" else 1.0
    views = [j.orbit_view(v0, step_deg * k) for k in range(args.warmup + steps)]
    r = reference_frames(j, verts, tris, w, h, views, warm=args.warmup, keep_last=cfg["shadow"])
    step_s = [c + s for c, s in zip(r["cast_s"], r["shade_s"])]
    total = sum(step_s)
    rays = w * h + (shadow_rays_of(r["pixels"]) if cfg["shadow"] else 0)
    value = rays * len(step_s) / total / 1e6
    metric = "primary+shadow Mrays/s @4K (ray cast + shading per frame)" if wl == "C" else "primary Mrays/s @1080p (ray cast + shading per frame)"
    sample = f"{tris.shape[0]}-triangle mesh: add_object (normals+bbox+QBVH) once, then {steps} frames {w}x{h} (cast + canvas_to_image), {args.warmup} warm-up frames discarded"
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(step_s),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(wl, f, tris.shape[0], w, h), "l2": "inputs_larger_than_l2"},
        "bvh_build_ms": 1e3 * r["build_s"], "cast_ms": 1e3 * statistics.median(r["cast_s"]), "shade_ms": 1e3 * statistics.median(r["shade_s"]),
        "cast_only_mrays_s": rays / statistics.median(r["cast_s"]) / 1e6,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": r["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# B200 arm — common set-up
# ----------------------------------------------------------------------------------------------
class Rig:
    """torch stream + NCCL group + library context of one rank."""

    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        import j3d_b200 as j
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
        self.torch, self.dist, self.j = torch, dist, j
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.numa_cpus = pin_to_gpu_numa(local_rank) if world > 1 else None
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.saved_stdout = None
        if world > 1:
            # NCCL prints its version banner on the C-level stdout at the first collective; the contract is ONE JSON
            # line on stdout, so fd 1 points at stderr until the line is printed
            sys.stdout.flush()
            self.saved_stdout = os.dup(1)
            os.dup2(2, 1)
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = j.Context(local_rank)
        # every kernel of the library, the NCCL collectives and the timing events share ONE stream
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.ctx.set_stream(self.stream.cuda_stream)
        mc, cav = j.make_matcap(0)
        self.ctx.set_matcap(mc, cav)
        # Frames in flight per GPU: more contexts of the same device, each with its own stream.  Frame k is rendered by
        # context k mod L; while the longest rays of frame k are still being finished by a few warps, the cast kernels of
        # the next frames move into the SM slots that frame k has already left (csrc/cast.cu, plain launch).  Meshes are
        # shared between the contexts of one device.
        self.lanes = max(1, min(4, args.lanes))
        self.ctxs, self.streams = [self.ctx], [self.stream]
        for _ in range(1, self.lanes):
            c2 = j.Context(local_rank)
            s2 = torch.cuda.Stream(device=self.dev)
            c2.set_stream(s2.cuda_stream)
            c2.set_matcap(mc, cav)
            self.ctxs.append(c2)
            self.streams.append(s2)
        # the multi-GPU plumbing lives behind the C ABI (j3dg_group_*, csrc/group.cu); torch.distributed only ships the NCCL id
        self.group = None
        if world > 1:
            from j3d_b200.dist import make_group
            self.group = make_group(self.ctx)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_over_ranks(self, *vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def build_and_broadcast(self, verts, tris, nv, nt, rebuilds=3):
        """BVH built once on rank 0; the other ranks receive nodes + triangle records over NCCL.
        Returns (mesh, info, build_ms (rank 0), mesh_create wall ms (rank 0), broadcast ms)."""
        torch, dist, ctx = self.torch, self.dist, self.ctx
        t0 = time.perf_counter()
        create_ms = bcast_ms = build_ms = None
        mesh = None
        if self.rank == 0:
            mesh = ctx.mesh_create(verts, tris)
            ctx.synchronize()
            create_ms = 1e3 * (time.perf_counter() - t0)  # host arrays -> usable BVH (H2D + build)
            b = [mesh.info().build_ms]
            for _ in range(rebuilds):
                mesh.rebuild()
                b.append(mesh.info().build_ms)
            build_ms = statistics.median(b[1:]) if rebuilds else b[0]
        if self.world > 1:
            # NCCL sets its channels up inside the first collective of each kind (hundreds of ms): warm them with a
            # one-triangle mesh so the timed broadcast measures the transfer
            tiny = None
            if self.rank == 0:
                tiny = ctx.mesh_create(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32))
            tiny = self.group.broadcast_mesh(tiny, root=0)
            ctx.synchronize()
            tiny.destroy()
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mesh = self.group.broadcast_mesh(mesh if self.rank == 0 else None, root=0)  # BVH + records + indexed geometry, NCCL over NVLink
            e1.record()
            torch.cuda.synchronize()
            bcast_ms = e0.elapsed_time(e1)
        return mesh, mesh.info(), build_ms, create_ms, bcast_ms

    def finish(self, line):
        if self.rank == 0:
            if self.saved_stdout is not None:
                sys.stdout.flush()
                os.dup2(self.saved_stdout, 1)
                os.close(self.saved_stdout)
            print(json.dumps(line), flush=True)
        if self.world > 1:
            if self.group is not None:
                self.group.destroy()
            self.dist.destroy_process_group()


def traffic_citation():
    """DRAM bytes per cast launch from the committed ncu capture (a citation of profiles/, not a per-run measurement)."""
    tp = ROOT / "profiles" / "traffic.json"
    try:
        d = json.loads(tp.read_text())
        return d.get("cast_kernel_dram_bytes_per_launch"), d.get("source", "profiles/traffic.json")
    except Exception:
        return None, None


# ----------------------------------------------------------------------------------------------
# workloads A / B: orbit sweep, frames sharded over the ranks
# ----------------------------------------------------------------------------------------------
def run_orbit(args, wl, rank, world, local_rank):
    rig = Rig(args, rank, world, local_rank)
    torch, dist, j, ctx, dev, stream = rig.torch, rig.dist, rig.j, rig.ctx, rig.dev, rig.stream
    cfg = WORKLOADS[wl]
    W, H = cfg["w"], cfg["h"]
    f = args.f or cfg["f"]
    verts, tris = j.icosphere(f)  # mesh replicated: every rank can generate it; only rank 0 uploads and builds
    nt, nv = tris.shape[0], verts.shape[0]
    mn, mx = j.compute_bb(verts)
    v0 = j.make_view(W, H, mn, mx)
    mesh, info, build_ms, create_ms, bcast_ms = rig.build_and_broadcast(verts, tris, nv, nt)

    ctxs, streams, L = rig.ctxs, rig.streams, rig.lanes
    pxs = [torch.empty((H, W, 32), dtype=torch.uint8, device=dev) for _ in range(L)]  # one canvas per frame in flight
    px = pxs[0]
    rgba2 = [torch.empty((H, W), dtype=torch.int32, device=dev) for _ in range(max(2, L))]
    # N > 1: every rank's frame must end up on rank 0 each step.
    #   --exchange peer (default): the shade kernel of every rank stores its RGBA straight into rank 0's HBM over
    #       NVLink peer memory (CUDA-IPC mapping, stream-ordered arrival / release flags; j3d_b200/dist.py::PeerFrames)
    #   --exchange nccl: dist.gather on a second stream, double-buffered RGBA (kept for comparison).
    pf = None
    comm = None
    if world > 1 and args.exchange == "peer":
        from j3d_b200.dist import PeerFrames
        try:
            # --slots-per-lane 2 lets a rank run further ahead of rank 0's consumption; measured on 8 GPUs it buys nothing
            # (20.9 vs 20.8 Grays/s, profiles/r2_n8_exchange_slots_ab.log): the ranks are not waiting for each other
            pf = PeerFrames(ctx, H, W, dev, dst=0, group=rig.group, nslots=args.slots_per_lane * max(2, L) if L > 1 else 2)
        except RuntimeError as e:  # every rank raises together: CUDA IPC is not available here, gather with NCCL instead
            if rank == 0:
                print(f"bench.py: {e}; using --exchange nccl", file=sys.stderr)
            args.exchange = "nccl"
    if world > 1 and args.exchange == "nccl":
        comm = torch.cuda.Stream(device=dev)
        L = 1  # the gather path keeps one frame in flight
    for sl in range(pf.nslots if (pf is not None and L > 1) else 0):
        pf.set_lane(sl, ctxs[sl % L])  # begin / arrive / release of the frames of slot sl go to the stream of the context that renders them
    gather_lists = [[torch.empty_like(rgba2[0]) for _ in range(world)] for _ in range(2)] if (comm is not None and rank == 0) else [None, None]
    ev_render = [torch.cuda.Event() for _ in range(2)]
    ev_gather = [torch.cuda.Event() for _ in range(2)]
    state = {"k": 0, "checksum": torch.zeros((), dtype=torch.int64, device=dev)}

    def step(v, lanes=None, local=False):
        lanes = L if lanes is None else lanes
        if pf is not None and not local:
            k = pf.begin()
            ln = k % L   # slot k mod nslots, nslots a multiple of L: the frames of a slot always run on the same lane
            ctxs[ln].render_frame([mesh], [], v, pixels_out=pxs[ln], rgba_out=pf.target(k))
            pf.arrive(k)
            # rank 0 consumes between arrival and release (PeerFrames protocol): nothing in the timed loop — the frames
            # are what the sweep produces; `verify_exchange` below consumes every frame of a short run instead
            pf.release(k)
            return
        k = state["k"]
        state["k"] = k + 1
        b = k & 1
        ln = k % lanes
        if comm is not None and k >= 2:
            stream.wait_event(ev_gather[b])  # the gather of frame k - 2 has left this buffer
        ctxs[ln].render_frame([mesh], [], v, pixels_out=pxs[ln], rgba_out=rgba2[b if comm is not None else ln])
        if comm is not None:
            ev_render[b].record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(ev_render[b])
                dist.gather(rgba2[b], gather_lists[b], dst=0)
                ev_gather[b].record(comm)

    def drain():
        if comm is not None:
            stream.wait_stream(comm)
        for s2 in streams[1:]:
            stream.wait_stream(s2)   # the timing events live on lane 0's stream

    def timed(views, warm, lanes=None, local=False):
        """Device-resident loop: `warm` untimed frames, then the rest between barriers; ms = max over ranks."""
        for v in views[:warm]:
            step(v, lanes, local)
        drain()
        rig.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s2 in streams[1:]:
            s2.wait_stream(stream)   # nothing of lane 1 starts before the first event
        for v in views[warm:]:
            step(v, lanes, local)
        drain()
        e1.record()
        rig.barrier()
        return rig.max_over_ranks(e0.elapsed_time(e1))[0]

    def timings_reset():
        tms = [c.timings(reset=True) for c in ctxs]
        n = max(1, sum(t.cast_count for t in tms))
        return (sum(t.cast_ms for t in tms) / n, sum(t.shade_ms for t in tms) / max(1, sum(t.shade_count for t in tms)),
                sum(t.kernel_launches for t in tms), sum(t.cast_count for t in tms))

    # ---- (1) weak arm: `steps` frames per rank, the N ranks together render an N-times finer sweep of the same arc ----
    total = args.warmup + args.steps
    arc = [j.orbit_view(v0, ((k * world + rank) / world) % 360.0) for k in range(total)]
    for v in arc[: args.warmup]:  # first-touch allocations outside everything
        step(v)
    drain()
    rig.barrier()
    # ---- (0) one frame at a time (N = 1): the same frames back to back on ONE stream.  This is the latency of a single
    #      j3dg_render_frame and the isolated duration of the cast kernel that the roofline block divides by. ----
    single = None
    if comm is None:
        nsingle = min(args.steps, 60)
        timings_reset()
        single_ms = timed(arc[: args.warmup + nsingle], args.warmup, lanes=1, local=True)
        s_cast, s_shade, _, _ = timings_reset()
        single = {"frames": nsingle, "ms_per_frame": single_ms / nsingle, "mrays_s": W * H * nsingle / single_ms / 1e3, "cast_ms": s_cast, "shade_ms": s_shade,
                  "note": "one frame in flight: frames back to back on one stream through one context" + (" (every rank its own frames, no exchange)" if world > 1 else "")}
    timings_reset()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(arc, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    cast_ms_inflight, shade_ms_inflight, nlaunch, ncast = timings_reset()
    rays_total = W * H * args.steps * world
    value = rays_total / (ms * 1e-3) / 1e6
    # the stage timers of overlapping frames include the time a kernel shares the machine with its neighbour: the
    # isolated figures come from the one-frame-at-a-time run when there is one
    cast_ms = single["cast_ms"] if single else cast_ms_inflight
    shade_ms = single["shade_ms"] if single else shade_ms_inflight
    launches = int(nlaunch * args.steps / max(1, ncast))  # launches of the timed frames of this rank

    # ---- exchange check: every frame of a short run, consumed between arrival and release, equals a local render ----
    exchange_ok = None
    if pf is not None:
        nchk = 6
        chk = [j.orbit_view(v0, 3.0 * (k * world + r)) for k in range(nchk) for r in range(world)]
        snaps = []
        for k in range(nchk):
            kk = pf.begin()
            ln = kk % L
            ctxs[ln].render_frame([mesh], [], chk[k * world + rank], pixels_out=pxs[ln], rgba_out=pf.target(kk))
            pf.arrive(kk)
            if rank == 0:
                with torch.cuda.stream(streams[ln]):
                    snaps.append(pf.frames(kk).clone())  # the consumer, on the stream that handed the frame over
            pf.release(kk)
        torch.cuda.synchronize()
        exchange_ok = all(c.status() == 0 for c in ctxs)
        if rank == 0:
            for k in range(nchk):
                for r in range(world):
                    ctx.render_frame([mesh], [], chk[k * world + r], pixels_out=px, rgba_out=rgba2[0])
                    torch.cuda.synchronize()
                    exchange_ok = exchange_ok and bool(torch.equal(snaps[k][r], rgba2[0]))
        rig.barrier()

    # ---- (2) strong arm (BASELINE configs[4]): the fixed 360-frame orbit, frame k on rank k mod N ----
    n360 = 360
    per = (n360 + world - 1) // world
    mine360 = [j.orbit_view(v0, float((k * world + rank) % 360)) for k in range(per)]
    sweep_ms = timed(mine360[:3] + mine360, 3)
    sweep360 = {"frames": n360, "scaling": "strong", "device": {"ms": sweep_ms, "frames_per_s": 1e3 * n360 / sweep_ms, "mrays_s": n360 * W * H / sweep_ms / 1e3}}

    # ---- (3) end to end through the C ABI with HOST buffers (pinned), D2H inside the timed region ----
    # The sweep call a user makes: j3dg_frame_submit / j3dg_frame_wait (pipelined j3dg_render_frame): the device->host
    # copy of frame k overlaps the kernels of frame k+1.  Every frame's host buffers are complete (frame_wait
    # returned) inside the timed region.  Each rank owns its frames' host buffers (frames sharded over the ranks).
    # host buffers: every context pipelines two frames of its own, so [lane][2]
    hpx = [[torch.empty((H, W, 32), dtype=torch.uint8).pin_memory() for _ in range(2)] for _ in range(L)]
    hrgba = [[torch.empty((H, W), dtype=torch.int32).pin_memory() for _ in range(2)] for _ in range(L)]
    centre = np.array([[W // 2, H // 2]], np.int32)
    last_buf = {"ln": 0, "b": 0}

    def sweep(vs, records):
        # frame k: context k mod L, host buffer (k div L) & 1 of that context; a context waits for its previous frame
        # right after it has submitted the next one (the copy of frame k - L overlaps the kernels of frames k - L + 1 .. k)
        n = len(vs)
        for k, v in enumerate(vs):
            ln, b = k % L, (k // L) & 1
            ctxs[ln].frame_submit([mesh], [], v, pixels_out=hpx[ln][b] if records else None, rgba_out=hrgba[ln][b])
            if k >= L:
                ctxs[ln].frame_wait()
        for k in range(max(0, n - L), n):
            ctxs[k % L].frame_wait()
        if vs:
            last_buf["ln"], last_buf["b"] = (n - 1) % L, ((n - 1) // L) & 1
            if not records:  # the records stayed in HBM: the host asks for the one under the cursor
                ctxs[(n - 1) % L].pick([mesh], [], vs[-1], centre)

    def timed_sweep(vs, warm, records):
        """Host wall clock around K = len(vs) - warm frames, max over ranks.  A short run (the driver's 20 steps last
        15 ms) is at the mercy of one host hiccup, so it is repeated (up to 5 times, every repetition its own warm-up and
        exactly K timed frames) and the MEDIAN is reported."""
        reps = max(1, min(5, 100 // max(1, len(vs) - warm)))
        times, nbytes = [], 0
        for _ in range(reps):
            sweep(vs[:warm], records)
            rig.barrier()
            for c in ctxs:
                c.readback_bytes(reset=True)
            t0 = time.perf_counter()
            sweep(vs[warm:], records)
            dt = time.perf_counter() - t0
            nbytes = sum(c.readback_bytes(reset=True) for c in ctxs) / max(1, len(vs) - warm)
            times.append(rig.max_over_ranks(dt)[0])
        return statistics.median(times), nbytes

    def set_dirty(on):
        for c in ctxs:
            c.set_dirty_rect(on)

    set_dirty(True)
    e2e_s, e2e_bytes = timed_sweep(arc, args.warmup, records=False)          # headline e2e: RGBA only + pick
    s360_s, s360_bytes = timed_sweep(mine360[:3] + mine360, 3, records=False)
    rec_s, rec_bytes = timed_sweep(arc, args.warmup, records=True)            # pixel records + RGBA, dirty rectangle
    dirty_px, dirty_rgba = hpx[last_buf["ln"]][last_buf["b"]].clone(), hrgba[last_buf["ln"]][last_buf["b"]].clone()
    set_dirty(False)
    full_s, full_bytes = timed_sweep(arc, args.warmup, records=True)          # every byte of both buffers, every frame
    dirty_identical = bool(torch.equal(dirty_px, hpx[last_buf["ln"]][last_buf["b"]]) and torch.equal(dirty_rgba, hrgba[last_buf["ln"]][last_buf["b"]]))
    sweep360["e2e"] = {"ms": 1e3 * s360_s, "frames_per_s": n360 / s360_s, "mrays_s": n360 * W * H / s360_s / 1e6, "d2h_bytes_per_frame": int(s360_bytes),
                       "api": "j3dg_frame_submit(pixels_out=NULL)/j3dg_frame_wait, RGBA to pinned host memory on the rendering rank"}
    # the same frames one by one through the synchronous j3dg_render_frame (kernels, then copy)
    for v in arc[:3]:  # untimed: the first call allocates the context's own canvas
        ctx.render_frame([mesh], [], v, pixels_out=hpx[0][0], rgba_out=hrgba[0][0])
    nsync = min(args.steps, 20)
    t0 = time.perf_counter()
    for v in arc[args.warmup: args.warmup + nsync]:
        ctx.render_frame([mesh], [], v, pixels_out=hpx[0][0], rgba_out=hrgba[0][0])
    sync_ms = 1e3 * (time.perf_counter() - t0) / nsync
    t0 = time.perf_counter()
    for v in arc[args.warmup: args.warmup + nsync]:
        ctx.render_frame([mesh], [], v, pixels_out=None, rgba_out=hrgba[0][0])
    sync_rgba_ms = 1e3 * (time.perf_counter() - t0) / nsync
    ctx.set_dirty_rect(True)   # the literal drop-in call with persistent host canvases: only the rectangle that changed crosses PCIe
    for v in arc[:2]:
        ctx.render_frame([mesh], [], v, pixels_out=hpx[0][0], rgba_out=hrgba[0][0])
    t0 = time.perf_counter()
    for v in arc[args.warmup: args.warmup + nsync]:
        ctx.render_frame([mesh], [], v, pixels_out=hpx[0][0], rgba_out=hrgba[0][0])
    sync_dirty_ms = 1e3 * (time.perf_counter() - t0) / nsync
    ctx.set_dirty_rect(False)
    h2d_bytes = ctypes.sizeof(j.View) + 256  # the view (kernel parameters) + the per-mesh table

    extras = None
    if world == 1:
        # ---- the other query clients of the same BVH, timed through the C ABI ----
        qxy = np.stack(np.meshgrid(np.arange(0, W, 8), np.arange(0, H, 8)), -1).reshape(-1, 2).astype(np.int32)
        ctx.pick([mesh], [], arc[-1], qxy)
        t0 = time.perf_counter()
        picks = ctx.pick([mesh], [], arc[-1], qxy)
        pick_ms = 1e3 * (time.perf_counter() - t0)
        vox_dim = 512
        grid = torch.empty((vox_dim ** 3 + 64,), dtype=torch.uint8, device=dev)
        mesh.voxelize(vox_dim, out=grid)
        t0 = time.perf_counter()
        mesh.voxelize(vox_dim, out=grid)
        vox_ms = 1e3 * (time.perf_counter() - t0)
        occupied = int((grid != 0).sum().item())
        del grid
        extras = {"pick": {"queries": int(qxy.shape[0]), "hits": int((picks["db_id"] != 0).sum()), "ms": pick_ms,
                           "note": "j3dg_pick on the resident canvas, host xy in / host records out (64 B per query)"},
                  "voxelize": {"max_dim": vox_dim, "rays": 3 * vox_dim * vox_dim, "occupied_voxels": occupied, "ms": vox_ms,
                               "mrays_s": 3 * vox_dim * vox_dim / vox_ms / 1e3,
                               "note": "j3dg_mesh_voxelize into a device grid: memset + 3 all-hits ray grids (vox.cpp:300-379)"}}

    # ---- roofline of the dominant kernel (cast): algorithmic bytes / live event time ----
    nodes_per_ray, tris_per_ray = ctx.cast_stats([mesh], arc[args.warmup])
    bytes_per_ray = nodes_per_ray * info.node_bytes + tris_per_ray * info.triangle_bytes + 32
    peak, peak_src = peaks()
    achieved = (W * H * bytes_per_ray) / (cast_ms * 1e-3) / 1e9
    traffic, traffic_src = traffic_citation() if world == 1 else (None, None)
    roofline = {"bound": "hbm", "kernel": "cast stage (cast_kernel: one persistent launch per ray type), timed alone — one frame in flight", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "bytes_per_ray": bytes_per_ray,
                "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray, "kernel_ms": cast_ms, "kernel_mrays_s": W * H / cast_ms / 1e3,
                "achieved_frames_in_flight": (W * H * bytes_per_ray) / (ms / args.steps * 1e-3) / 1e9 if world == 1 else None,
                "note": "achieved = ALGORITHMIC bytes (SURVEY §8d) over the live CUDA-event stage time; L1/L2 absorb re-use, so this is not a DRAM-bandwidth share (traffic = DRAM bytes per launch from the cited ncu capture)"}

    # ---- BASELINE configs[2] on the same box(es): 300 M triangles, 3840x2160 + shadow rays, ONE frame sharded over the
    #      N ranks (collective: every rank takes part) ----
    cstage = None
    if not args.no_config_c:
        if pf is not None:
            pf.close()
            pf = None
        cm = sharded_frame_measure(rig, args, steps=6, warmup=2)
        if rank == 0:
            cstage = {"workload": workload_name("C", cm["f"], cm["nt"], cm["W"], cm["H"]), "n_gpus": world, "frames": cm["steps"], "ms_per_frame": cm["ms"] / cm["steps"],
                      "mrays_s": cm["rays"] / cm["ms"] / 1e3, "rays_per_frame": cm["rays"] / cm["steps"], "frames_per_s": 1e3 * cm["steps"] / cm["ms"],
                      "cast_ms_max_rank": cm["cast_ms"], "shade_ms_max_rank": cm["shade_ms"], "bvh_build_ms": cm["build_ms"], "bvh_bytes": cm["bvh_bytes"],
                      "bvh_broadcast_ms": cm["bcast_ms"], "mesh_create_ms": cm["create_ms"], "mesh_generate_s": cm["gen_s"],
                      "build_roofline_frac": (12 * cm["nt"] + 12 * cm["nv"] + cm["bvh_bytes"]) / (cm["build_ms"] * 1e-3) / 1e9 / cm["peak"] if cm["build_ms"] else None,
                      "sharded_frame_equals_unsharded": cm["identical"], "hit_pixels": cm["hit"], "shadowed_pixels": cm["shadowed"], "exchange_status": cm["status"],
                      "sharding": "one GPU" if world == 1 else f"32-row screen bands round-robin over {world} ranks inside the cast / shade kernels, BVH NCCL-broadcast, bands stored into ONE frame in rank 0's HBM over NVLink peer memory"}

    if rank != 0:
        if pf is not None:
            pf.close()
        rig.finish(None)
        return 0

    # ---- CPU baseline (the real reference on this box's host cores) + parity of the same frame, N = 1 only ----
    cpu = parity = None
    bvh_bytes = int(info.nr_of_nodes) * info.node_bytes + nt * info.triangle_bytes
    stages = {"build": {"ms": build_ms, "algorithmic_bytes": 12 * nt + 12 * nv + bvh_bytes, "mtris_s": nt / build_ms / 1e3 if build_ms else None,
                        "roofline_frac": (12 * nt + 12 * nv + bvh_bytes) / (build_ms * 1e-3) / 1e9 / peak if build_ms else None},
              "shade": {"ms": shade_ms, "algorithmic_bytes": 40 * W * H, "roofline_frac": 40 * W * H / (shade_ms * 1e-3) / 1e9 / peak if shade_ms else None}}
    if world == 1 and not args.no_cpu_baseline:
        from oracle.bindings import ref_available
        if ref_available():
            from parity import parity_stats
            nref = 3
            r = reference_frames(j, verts, tris, W, H, arc[args.warmup - 1: args.warmup + nref], warm=1, keep_last=True)
            t_all = [c + s for c, s in zip(r["cast_s"], r["shade_s"])]
            cpu = {"value": W * H * len(t_all) / sum(t_all) / 1e6, "unit": "Mrays/s", "cores": r["cores"], "kind": "reference",
                   "bvh_build_ms": 1e3 * r["build_s"], "ms_per_frame": 1e3 * sum(t_all) / len(t_all),
                   "sample": f"same mesh ({nt} triangles): add_object once + {nref} frames {W}x{H} of the timed sweep after 1 warm-up frame"}
            # parity on the same run: the reference's last frame against the GPU's frame of the same view
            got_px = np.zeros((H, W), j.PIXEL_DTYPE)
            got_rgba = np.zeros((H, W), np.uint32)
            ctx.render_frame([mesh], [], arc[args.warmup + nref - 1], pixels_out=got_px, rgba_out=got_rgba)
            parity = parity_stats(got_px, r["pixels"], got_rgba, r["image"])
            parity["frame"] = f"step {nref - 1} of the timed sweep, reference pixel buffer + image of the cpu_baseline leg"
            parity["pass"] = bool(parity["id_agree_frac"] >= 0.9999 and parity["unclassified"] == 0 and parity.get("max_rel_depth", 0.0) <= 1e-5
                                  and parity.get("max_abs_bary", 0.0) <= 1e-5 and parity["rgba_within_1lsb_frac"] >= 0.999)
        else:
            cpu = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
        # ---- splat stage (BASELINE configs[3]) on the same box ----
        if not args.no_splat:
            stages["splat"] = splat_stage(rig, args.points or WORKLOADS["P"]["points"], peak, with_cpu=True)
        if not args.no_ingest:
            stages["ingest"] = ingest_stage(rig, verts, tris, peak, with_cpu=True)
    del verts, tris

    line = {
        "metric": "primary Mrays/s @1080p (ray cast + shading per frame)", "value": value, "unit": "Mrays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(wl, f, nt, W, H), "l2": "inputs_larger_than_l2"},  # the same two keys as the reference arm's line
        "frames_in_flight": {"per_gpu": L, "how": f"frame k is rendered by context k mod {L} of the same device ({L} streams, meshes shared): the cast kernels of the next frames move into the SM slots frame k has left while a few warps still finish its longest rays; every frame is complete and in its output buffer inside the timed region" if L > 1 else "one context, frames back to back on one stream",
                             "one_frame_at_a_time": single},
        "layout": {"poses": f"frame i at i/{world} degrees on rank i mod {world}: every N renders the same arc, {world}x finer" if world > 1 else "frame i at i degrees",
                   "sharding": "replicas only" if world == 1 else f"orbit frames round-robin over {world} ranks, mesh + BVH built on rank 0 and replicated by j3dg_group_broadcast_mesh (NCCL), " + ("every rank's shade kernel stores its RGBA into rank 0's HBM over NVLink peer memory (CUDA IPC), stream-ordered arrival/release flags" if args.exchange == "peer" else "RGBA NCCL-gathered on rank 0 every step (second stream, overlapping the kernels of frame k+1)"),
                   "bvh_bytes": bvh_bytes},
        "bvh_build_ms": build_ms, "bvh_nodes": int(info.nr_of_nodes), "frames_per_s": 1e3 * args.steps * world / ms,
        "cast_ms": cast_ms, "shade_ms": shade_ms,
        "e2e": {"value": rays_total / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(e2e_bytes),
                "ms_per_step": 1e3 * e2e_s / args.steps, "repetitions_median_of": max(1, min(5, 100 // max(1, args.steps))),
                "api": "j3dg_frame_submit(pixels_out=NULL)/j3dg_frame_wait (pipelined, pinned persistent host buffer, dirty-rectangle readback) + j3dg_pick: RGBA crosses PCIe every frame, the pixel records stay in HBM and are queried on demand",
                "sync_render_frame_rgba_ms_per_step": sync_rgba_ms, "mesh_create_ms": create_ms},
        "e2e_full_records": {"value": rays_total / rec_s / 1e6, "unit": "Mrays/s", "ms_per_step": 1e3 * rec_s / args.steps, "d2h_bytes_per_step": int(rec_bytes),
                             "api": "j3dg_frame_submit/j3dg_frame_wait with pixels_out AND rgba_out (36 B/pixel to the host, dirty rectangle)",
                             "host_buffers_identical_to_full_copy": dirty_identical,
                             "full_copy": {"value": rays_total / full_s / 1e6, "ms_per_step": 1e3 * full_s / args.steps, "d2h_bytes_per_step": int(full_bytes)},
                             "sync_render_frame_ms_per_step": sync_ms, "sync_render_frame_dirty_rect_ms_per_step": sync_dirty_ms},
        "sweep360": sweep360,
        "gpu_launches": launches,
        "clocks": clocks, "roofline": roofline, "stages": stages,
    }
    if rig.numa_cpus is not None:
        line["layout"]["host_affinity"] = f"each rank pinned to the {rig.numa_cpus} CPUs next to its GPU (NVML affinity) before allocating pinned buffers"
    if cstage is not None:
        line["stages"]["config_c"] = cstage
    if extras is not None:
        line["queries"] = extras
    if bcast_ms is not None:
        line["bvh_broadcast_ms"] = bcast_ms
        line["exchange"] = {"kind": args.exchange, "verified": exchange_ok, "bytes_per_step_into_rank0": (world - 1) * W * H * 4}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if parity is not None:
        line["parity"] = parity
    if pf is not None:
        pf.close()
    rig.finish(line)
    return 0


# ----------------------------------------------------------------------------------------------
# splat (configs[3]) — stage of the default line at N = 1, and workload P
# ----------------------------------------------------------------------------------------------
def splat_stage(rig, n_points, peak, with_cpu, reps=5):
    torch, j, ctx, dev = rig.torch, rig.j, rig.ctx, rig.dev
    W, H = WORKLOADS["P"]["w"], WORKLOADS["P"]["h"]
    pos, nrm, clr = j.cloud(n_points)
    mn, mx = j.compute_bb(pos)
    v0 = j.make_view(W, H, mn, mx)
    t0 = time.perf_counter()
    cl = ctx.cloud_create(pos, nrm, clr)
    ctx.synchronize()
    upload_ms = 1e3 * (time.perf_counter() - t0)
    px = torch.empty((H, W, 32), dtype=torch.uint8, device=dev)
    rgba = torch.empty((H, W), dtype=torch.int32, device=dev)
    for k in range(2):
        ctx.render_frame([], [cl], j.orbit_view(v0, float(k)), pixels_out=px, rgba_out=rgba)
    ctx.timings(reset=True)
    for k in range(reps):
        ctx.render_frame([], [cl], j.orbit_view(v0, float(2 + k)), pixels_out=px, rgba_out=rgba)
    tm = ctx.timings(reset=True)
    ms = tm.splat_ms / max(1, tm.splat_count)
    covered = int((px.view(torch.int32)[..., 7] != 0).sum().item())
    out = {"points": n_points, "ms": ms, "mpoints_s": n_points / ms / 1e3, "algorithmic_bytes": SPLAT_BYTES_PER_POINT * n_points + 44 * W * H,
           "roofline_frac": (SPLAT_BYTES_PER_POINT * n_points + 44 * W * H) / (ms * 1e-3) / 1e9 / peak, "pixels_covered": covered, "upload_ms": upload_ms,
           "workload": f"{n_points}-point vertex-coloured cloud (normals + colours), {W}x{H}, empty mesh scene, orbit 1 deg/frame"}
    cl.destroy()
    if with_cpu:
        from oracle.bindings import Ref, ref_available
        if ref_available():
            ns = min(n_points, 10_000_000)
            ref = Ref(W, H)
            ref.add_cloud(pos[:ns].copy(), nrm[:ns].copy(), clr[:ns].copy())
            ref.set_view(v0)
            ts = []
            for k in range(3):
                ref.set_view(j.orbit_view(v0, float(k)))
                ref.render(7)
                if k:
                    ts.append(ref.times()["splat"])
            out["cpu_reference"] = {"points": ns, "ms": 1e3 * sum(ts) / len(ts), "mpoints_s": ns * len(ts) / sum(ts) / 1e6, "cores": ref.cores(),
                                    "sample": f"first {ns} points of the same cloud, render_pointclouds_on_image, 2 frames after 1 warm-up"}
            ref.close()
    return out


def ply_bytes(verts, tris):
    """A binary little-endian PLY file of an indexed triangle mesh (float x y z; uchar count + 3 int indices per face)."""
    head = (f"ply\nformat binary_little_endian 1.0\ncomment bench.py\nelement vertex {verts.shape[0]}\nproperty float x\nproperty float y\n"
            f"property float z\nelement face {tris.shape[0]}\nproperty list uchar int vertex_indices\nend_header\n").encode("ascii")
    faces = np.empty(tris.shape[0], dtype=[("n", "u1"), ("i", "<i4", 3)])
    faces["n"] = 3
    faces["i"] = tris
    return head + np.ascontiguousarray(verts, "<f4").tobytes() + faces.tobytes()


def ingest_stage(rig, verts, tris, peak, with_cpu):
    """SURVEY §8f rank 4 on the same box: the config-B mesh as a binary PLY file decoded on the device (j3dg_ply_decode,
    file bytes in pinned host memory -> vertex / index arrays in HBM -> j3dg_mesh_create_from_ply), and k-NN normal
    estimation of a point cloud (j3dg_cloud_knn_normals / j3dg_cloud_estimate_normals), each with the reference's CPU code
    (jtk::read_ply_from_memory, estimate_normals) on a bounded sample."""
    torch, j, ctx = rig.torch, rig.j, rig.ctx
    data = ply_bytes(verts, tris)
    pinned = torch.empty(len(data), dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = np.frombuffer(data, np.uint8)
    nv, nt = verts.shape[0], tris.shape[0]
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        ply = ctx.ply_decode(pinned.numpy())
        wall = 1e3 * (time.perf_counter() - t0)
        info = ply.info()
        if best is None or wall < best[0]:
            best = (wall, info.upload_ms, info.decode_ms)
        ok = bool(info.nr_of_vertices == nv and info.nr_of_faces == nt)
        if _ < 2:
            ply.destroy()
    t0 = time.perf_counter()
    m = ply.to_mesh()
    ctx.synchronize()
    to_mesh_ms = 1e3 * (time.perf_counter() - t0)
    build_ms = m.info().build_ms
    m.destroy()
    ply.destroy()
    alg = len(data) + 12 * nv + 12 * nt
    out = {"ply": {"file_bytes": len(data), "vertices": nv, "faces": nt, "decode_call_ms": best[0], "upload_ms": best[1], "decode_kernels_ms": best[2],
                   "decode_gb_s": alg / (best[2] * 1e-3) / 1e9, "roofline_frac": alg / (best[2] * 1e-3) / 1e9 / peak, "algorithmic_bytes": alg,
                   "mesh_create_from_ply_ms": to_mesh_ms, "bvh_build_ms": build_ms, "counts_ok": ok,
                   "note": "decode = two kernels over the uploaded element data (file read once, records staged in shared memory); upload = one H2D copy of the file from pinned memory"}}
    del pinned
    n_pts, k = 2_000_000, 10
    pos, _, _ = j.cloud(n_pts)
    cl = ctx.cloud_create(pos)
    cl.knn_normals(k)  # warm-up (allocations)
    t0 = time.perf_counter()
    cl.knn_normals(k)
    knn_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    cl.estimate_normals(k)
    est_ms = 1e3 * (time.perf_counter() - t0)
    cl.destroy()
    out["normals"] = {"points": n_pts, "k": k, "knn_fit_ms": knn_ms, "mpoints_s": n_pts / knn_ms / 1e3, "estimate_normals_ms": est_ms,
                      "note": "knn_fit = grid build + k-NN + plane fit on the device, lists and normals copied to the host; estimate_normals adds the serial orientation walk on the host"}
    if with_cpu:
        from oracle.bindings import ref_available, ref_estimate_normals, ref_read_ply
        if ref_available():
            fs = min(nt, 2_000_000)
            sample = ply_bytes(verts, tris[:fs])
            t0 = time.perf_counter()
            got = ref_read_ply(sample)
            dt = time.perf_counter() - t0
            out["ply"]["cpu_reference"] = {"file_bytes": len(sample), "faces": fs, "ms": 1e3 * dt, "gb_s": (len(sample) + 12 * nv + 12 * fs) / dt / 1e9, "cores": 1,
                                           "ok": bool(got is not None and got["triangles"].shape[0] == fs),
                                           "sample": f"the same vertices + the first {fs} faces, jtk::read_ply_from_memory (rply, one thread; includes the binding's copy of the arrays)"}
            ns = 30_000  # the reference's estimate_normals is far from linear in practice: 200 000 points take 43 s on this box
            t0 = time.perf_counter()
            ref_estimate_normals(pos[:ns].copy(), k)
            dt = time.perf_counter() - t0
            out["normals"]["cpu_reference"] = {"points": ns, "ms": 1e3 * dt, "mpoints_s": ns / dt / 1e6, "cores": 1,
                                               "sample": f"first {ns} points, estimate_normals (j3d/pc.cpp:256: k-d tree, one thread)"}
    return out


def run_splat(args, rank, world, local_rank):
    """Workload P: replicas only (every rank splats the same cloud for its own frames)."""
    rig = Rig(args, rank, world, local_rank)
    peak, peak_src = peaks()
    n = args.points or WORKLOADS["P"]["points"]
    st = splat_stage(rig, n, peak, with_cpu=(rank == 0 and not args.no_cpu_baseline), reps=max(args.steps, 3))
    ms = rig.max_over_ranks(st["ms"])[0]
    W, H = WORKLOADS["P"]["w"], WORKLOADS["P"]["h"]
    line = None
    if rank == 0:
        line = {"metric": "Mpoints/s splatted @1080p", "value": world * n / ms / 1e3, "unit": "Mpoints/s", "n_gpus": world, "steps": max(args.steps, 3), "warmup": 2,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": st["workload"], "l2": "inputs_larger_than_l2"}, "layout": {"sharding": "replicas only"},
                "roofline": {"bound": "hbm", "kernel": "splat stage (seed + project/atomicMax + resolve)", "achieved": st["algorithmic_bytes"] / (ms * 1e-3) / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": st["algorithmic_bytes"] / (ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src},
                "stages": {"splat": st}, "gpu_launches": 5 * max(args.steps, 3),
                "e2e": {"value": None, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "stage bench: the frame's e2e is workload B's"}}
        if "cpu_reference" in st:
            c = st["cpu_reference"]
            line["cpu_baseline"] = {"value": c["mpoints_s"], "unit": "Mpoints/s", "cores": c["cores"], "kind": "reference", "sample": c["sample"]}
    rig.finish(line)
    return 0


# ----------------------------------------------------------------------------------------------
# workload C (configs[2]): ONE huge frame, screen bands sharded over the ranks into ONE frame in rank 0's HBM
# ----------------------------------------------------------------------------------------------
def sharded_frame_measure(rig, args, steps, warmup):
    """BASELINE configs[2]: ONE frame with shadow rays rendered by all ranks together (32-row screen bands round-robin over
    the ranks inside the cast / shade kernels, the bands stored into one frame in rank 0's HBM over NVLink peer memory).
    Collective: every rank calls it.  Returns the measurements (a dict; meaningful on rank 0)."""
    torch, dist, j, ctx, dev = rig.torch, rig.dist, rig.j, rig.ctx, rig.dev
    rank, world, local_rank = rig.rank, rig.world, rig.local_rank
    cfg = WORKLOADS["C"]
    W, H = cfg["w"], cfg["h"]
    f = args.f or cfg["f"]
    nt, nv = 20 * f * f, 10 * f * f + 2
    bb = torch.zeros(6, dtype=torch.float32, device=dev)
    verts = tris = None
    gen_s = None
    if rank == 0:  # only rank 0 needs the indexed mesh: the others receive the finished BVH (triangle records hold the vertices)
        t0 = time.perf_counter()
        verts, tris = j.icosphere(f)
        gen_s = time.perf_counter() - t0
        mn, mx = j.compute_bb(verts)
        bb = torch.tensor(np.concatenate([mn, mx]), dtype=torch.float32, device=dev)
    mesh, info, build_ms, create_ms, bcast_ms = rig.build_and_broadcast(verts, tris, nv, nt, rebuilds=1)
    del verts, tris
    if world > 1:
        dist.broadcast(bb, 0)
    bbh = bb.cpu().numpy()
    v0 = j.make_view(W, H, bbh[:3], bbh[3:], j.DEFAULT_FLAGS | j.SHADOW)
    total = warmup + steps
    views = [j.orbit_view(v0, 25.0 * k) for k in range(total)]
    px = torch.zeros((H, W, 32), dtype=torch.uint8, device=dev)
    rgba = torch.zeros((H, W), dtype=torch.int32, device=dev)
    pf = None
    if world > 1:
        from j3d_b200.dist import PeerFrames
        pf = PeerFrames(ctx, H, W, dev, dst=0, shared_frame=True, group=rig.group)
    ctx.set_screen_shard(rank, world)
    last = {}

    def frame(v):
        if pf is None:
            ctx.render_frame([mesh], [], v, pixels_out=px, rgba_out=rgba)
            return
        k = pf.begin()
        ctx.render_frame([mesh], [], v, pixels_out=px, rgba_out=pf.target(k))
        pf.arrive(k)
        last["k"] = k
        pf.release(k)

    for v in views[:warmup]:
        frame(v)
    rig.barrier()
    ctx.timings(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for v in views[warmup:]:
        frame(v)
    e1.record()
    rig.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = rig.max_over_ranks(e0.elapsed_time(e1))[0]
    tm = ctx.timings(reset=True)
    rays = rig.sum_over_ranks(float(tm.rays))[0]
    cast_ms, shade_ms = rig.max_over_ranks(tm.cast_ms / max(1, tm.cast_count), tm.shade_ms / max(1, tm.shade_count))
    launches = int(tm.kernel_launches)
    # parity of the sharded path at full size: rank 0 renders the last frame alone, unsharded, and compares
    identical = None
    hit = shadowed = None
    if rank == 0:
        got = (pf.frames(last["k"])[0] if pf is not None else rgba).clone()
        ctx.set_screen_shard(0, 1)
        ctx.render_frame([mesh], [], views[-1], pixels_out=px, rgba_out=rgba)
        torch.cuda.synchronize()
        identical = bool(torch.equal(got, rgba))
        p = px.cpu().numpy().view(j.PIXEL_DTYPE).reshape(H, W)
        hm = p["object_id"] != 0xFFFFFFFF
        hit, shadowed = int(hm.sum()), int((p["mark"][hm] & 1).sum())
    rig.barrier()
    ctx.set_screen_shard(0, 1)
    peak, _ = peaks()
    bvh_bytes = int(info.nr_of_nodes) * info.node_bytes + nt * info.triangle_bytes
    out = dict(f=f, nt=nt, nv=nv, W=W, H=H, ms=ms, rays=rays, cast_ms=cast_ms, shade_ms=shade_ms, launches=launches, identical=identical, hit=hit,
               shadowed=shadowed, build_ms=build_ms, create_ms=create_ms, bcast_ms=bcast_ms, gen_s=gen_s, info_nodes=int(info.nr_of_nodes), bvh_bytes=bvh_bytes,
               clocks=clocks, status=ctx.status(), peak=peak, steps=steps, warmup=warmup)
    if pf is not None:
        pf.close()
    mesh.destroy()
    return out


def run_sharded_frame(args, rank, world, local_rank):
    rig = Rig(args, rank, world, local_rank)
    m = sharded_frame_measure(rig, args, args.steps, args.warmup)
    line = None
    if rank == 0:
        value = m["rays"] / m["ms"] / 1e3
        line = {"metric": "primary+shadow Mrays/s @4K (ray cast + shading per frame)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": m["ms"] / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": workload_name("C", m["f"], m["nt"], m["W"], m["H"]), "l2": "inputs_larger_than_l2"},
                "layout": {"sharding": "replicas only" if world == 1 else f"32-row screen bands round-robin over {world} ranks inside the cast / shade kernels, BVH NCCL-broadcast, every rank's shade kernel stores its bands into ONE frame in rank 0's HBM over NVLink peer memory",
                           "bvh_bytes": m["bvh_bytes"]},
                "frames_per_s": 1e3 * args.steps / m["ms"], "rays_per_frame": m["rays"] / args.steps, "cast_ms_max_rank": m["cast_ms"], "shade_ms_max_rank": m["shade_ms"],
                "bvh_build_ms": m["build_ms"], "bvh_nodes": m["info_nodes"], "mesh_create_ms": m["create_ms"], "bvh_broadcast_ms": m["bcast_ms"], "mesh_generate_s": m["gen_s"],
                "sharded_frame_equals_unsharded": m["identical"], "hit_pixels": m["hit"], "shadowed_pixels": m["shadowed"],
                "stages": {"build": {"ms": m["build_ms"], "algorithmic_bytes": 12 * m["nt"] + 12 * m["nv"] + m["bvh_bytes"],
                                     "roofline_frac": (12 * m["nt"] + 12 * m["nv"] + m["bvh_bytes"]) / (m["build_ms"] * 1e-3) / 1e9 / m["peak"] if m["build_ms"] else None}},
                "gpu_launches": m["launches"], "clocks": m["clocks"], "exchange_status": m["status"],
                "e2e": {"value": None, "unit": "Mrays/s", "h2d_bytes_per_step": ctypes.sizeof(rig.j.View) + 256, "d2h_bytes_per_step": 0,
                        "note": "the frame stays in rank 0's HBM; host readback is measured on workload B"}}
    rig.finish(line)
    return 0


def main():
    # torchrun pins OMP_NUM_THREADS to 1; the procedural mesh generator (libj3d_synth, OpenMP) would then build the
    # 300 M-triangle mesh of config C on one core (21 s instead of 1.4 s).  Set before the library is loaded.
    if os.environ.get("OMP_NUM_THREADS") == "1" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ["OMP_NUM_THREADS"] = str(max(1, min(16, (os.cpu_count() or 16) // int(os.environ["WORLD_SIZE"]))))
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="B", choices=sorted(WORKLOADS))
    ap.add_argument("--f", type=int, default=0, help="override the icosphere frequency (T = 20 f^2)")
    ap.add_argument("--points", type=int, default=0, help="override the number of points of the splat stage / workload P")
    ap.add_argument("--slots-per-lane", type=int, default=1, help="N > 1: slots of the frame exchange per lane")
    ap.add_argument("--lanes", type=int, default=3, help="frames in flight per GPU (1 .. 4 contexts / streams)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-splat", action="store_true", help="skip the splat stage of the default line")
    ap.add_argument("--no-ingest", action="store_true", help="skip the PLY decode / normal estimation stage of the default line")
    ap.add_argument("--no-config-c", action="store_true", help="skip the 300 M-triangle sharded 4K frame (BASELINE configs[2]) of the default line")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="N > 1: how every rank's RGBA frame reaches rank 0")
    args = ap.parse_args()
    wl = args.workload
    if not args.steps:
        args.steps = {"A": 360, "B": 360, "C": 8, "P": 5}[wl]
    if not args.warmup:
        args.warmup = {"A": 10, "B": 10, "C": 3, "P": 3}[wl]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        return run_reference(args, wl, rank, world)
    from j3d_b200 import build
    if not (ROOT / "j3d_b200" / "libj3dg.so").exists():
        build.build_all()
    if wl == "P":
        return run_splat(args, rank, world, local_rank)
    if wl == "C":
        return run_sharded_frame(args, rank, world, local_rank)
    return run_orbit(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
