#!/usr/bin/env python
"""bench.py — headline benchmark of the j3d hot path on B200 (contract: see DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload B|A] [--f F]

Workload (BASELINE.json configs[1]): synthetic 28 037 120-triangle noised geodesic icosphere
("Lucy scale", f = 1184), 1920x1080, default settings (edges + matcap shading), default camera +
unzoom pose.  A *step* is one frame of the hot path: ray cast (one primary ray per pixel, misses
included) + fused shading, the camera orbiting 1 degree per step so successive frames touch
different parts of the 1.6 GB BVH (inputs larger than the 126 MB L2; nothing is cached between
steps).  metric = primary Mrays/s = W*H*steps / time.  The BVH build is timed separately
(`bvh_build_ms`, CUDA events around the build kernels, median of 3 rebuilds from resident data).

N > 1 (torchrun, one rank per GPU): the 360-frame orbit sweep is sharded frame-wise (BASELINE
configs[4]) — rank r renders frames r, r+N, ...; the BVH is built once on rank 0 and broadcast
with NCCL; every step ends with an NCCL gather of the RGBA frame on rank 0.  Weak scaling: each
rank renders `steps` frames.

--impl reference: the reference's own std::thread CPU renderer (oracle/_ref, compiled from the
unmodified j3d sources) on the same mesh / camera path, on this box's host cores.
"""
from __future__ import annotations

import argparse
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

W, H = 1920, 1080
WORKLOADS = {"B": 1184, "A": 59}


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


def make_mesh(f):
    import j3d_b200 as j
    return j.icosphere(f)


def frame_views(j, v0, first, count, stride):
    return [j.orbit_view(v0, float((first + k * stride) % 360)) for k in range(count)]


# ----------------------------------------------------------------------------------------------
# reference arm: the unmodified reference renderer on the host cores
# ----------------------------------------------------------------------------------------------
def run_reference(args, f, rank, world):
    if rank != 0:
        return 0
    import j3d_b200 as j
    from oracle.bindings import Ref, ref_available
    kind = "reference"
    if not ref_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libj3d_ref.so was not built (reference sources absent at build time)"}))
        return 0
    verts, tris = make_mesh(f)
    ref = Ref(W, H)
    cores = ref.cores()
    ref.add_mesh(verts, tris)
    build_s = ref.times()["build"]
    ref.unzoom()
    v0 = ref.view()
    views = frame_views(j, v0, 0, args.warmup + args.steps, 1)
    cast, shade = [], []
    for k, v in enumerate(views):
        ref.set_view(v)
        ref.render(3)  # cast + shade (no point clouds in this workload)
        t = ref.times()
        if k >= args.warmup:
            cast.append(t["cast"]); shade.append(t["shade"])
    step_s = [c + s for c, s in zip(cast, shade)]
    total = sum(step_s)
    value = W * H * len(step_s) / total / 1e6
    sample = f"{tris.shape[0]}-triangle mesh: add_object (normals+bbox+QBVH) once, then {args.steps} frames {W}x{H} (cast + canvas_to_image), {args.warmup} warm-up frames discarded"
    line = {
        "impl": "reference", "metric": "primary Mrays/s @1080p (ray cast + shading per frame)", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(step_s),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(f, tris.shape[0]), "l2": "inputs_larger_than_l2"},
        "bvh_build_ms": 1e3 * build_s, "cast_ms": 1e3 * statistics.median(cast), "shade_ms": 1e3 * statistics.median(shade),
        "cast_only_mrays_s": W * H / statistics.median(cast) / 1e6,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_name(f, nt):
    return f"noised geodesic icosphere f={f} ({nt} triangles), {W}x{H}, 1 primary ray/pixel + fused shading, orbit 1 deg/step"


def cpu_baseline(j, f, verts, tris, v0, budget_frames=3):
    """Bounded CPU sample of the same workload through the real reference (rank 0, N=1)."""
    from oracle.bindings import Ref, ref_available
    if not ref_available():
        return {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    ref = Ref(W, H)
    ref.add_mesh(verts, tris)
    build_s = ref.times()["build"]
    ref.unzoom()
    t_all = []
    for k, v in enumerate(frame_views(j, v0, 0, 1 + budget_frames, 1)):
        ref.set_view(v)
        ref.render(3)
        t = ref.times()
        if k >= 1:
            t_all.append(t["cast"] + t["shade"])
    cores = ref.cores()
    ref.close()
    return {"value": W * H * len(t_all) / sum(t_all) / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference",
            "bvh_build_ms": 1e3 * build_s, "ms_per_frame": 1e3 * sum(t_all) / len(t_all),
            "sample": f"same mesh ({tris.shape[0]} triangles): add_object once + {budget_frames} frames {W}x{H} after 1 warm-up frame"}


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def run_b200(args, f, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import j3d_b200 as j

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on the C-level stdout at the first collective; the contract is ONE JSON line
        # on stdout, so fd 1 points at stderr until the line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    ctx = j.Context(local_rank)
    # every kernel of the library, the NCCL collectives and the timing events share ONE stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    verts, tris = make_mesh(f)  # mesh replicated: every rank holds the full geometry in HBM
    nt = tris.shape[0]
    mn, mx = j.compute_bb(verts)
    v0 = j.make_view(W, H, mn, mx)
    mc, cav = j.make_matcap(0)
    ctx.set_matcap(mc, cav)

    # ---- BVH: built once on rank 0; other ranks receive nodes + triangle records over NCCL ----
    t0 = time.perf_counter()
    e2e_build_ms = None
    bcast_ms = None
    if world == 1 or rank == 0:
        mesh = ctx.mesh_create(verts, tris)
        ctx.synchronize()
        e2e_build_ms = 1e3 * (time.perf_counter() - t0)  # host arrays -> usable BVH (H2D + build)
    if world > 1:
        from j3d_b200.dist import broadcast_bvh
        meta = torch.zeros(4, dtype=torch.int64, device=dev)
        if rank == 0:
            meta[0] = mesh.info().nr_of_nodes
        dist.broadcast(meta, 0)
        if rank != 0:
            mesh = ctx.mesh_create_empty(verts.shape[0], nt, int(meta[0].item()))
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        broadcast_bvh(mesh, src=0, device=dev)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    info = mesh.info()
    build_ms = None
    if rank == 0:
        b = []
        for _ in range(3):
            mesh.rebuild()
            b.append(mesh.info().build_ms)
        build_ms = statistics.median(b)
        if world > 1:  # keep every rank on the same tree
            dist.barrier()
    elif world > 1:
        dist.barrier()
    if world > 1:
        from j3d_b200.dist import broadcast_bvh
        broadcast_bvh(mesh, src=0, device=dev)

    # ---- device-resident frame buffers ----
    px = torch.empty((H, W, 32), dtype=torch.uint8, device=dev)
    rgba2 = [torch.empty((H, W), dtype=torch.int32, device=dev) for _ in range(2)]
    rgba = rgba2[0]
    # N > 1: every rank's frame must end up on rank 0 each step.
    #   --exchange peer (default): the shade kernel of every rank stores its RGBA straight into rank 0's HBM over
    #       NVLink peer memory (CUDA-IPC mapping, stream-ordered arrival / release flags; j3d_b200/dist.py::PeerFrames)
    #       — no collective kernel has to find room beside the cooperative cast kernel, which owns every SM.
    #   --exchange nccl: dist.gather on a second stream, double-buffered RGBA.
    comm = torch.cuda.Stream(device=dev) if (world > 1 and args.exchange == "nccl") else None
    gather_lists = [[torch.empty_like(rgba) for _ in range(world)] for _ in range(2)] if (comm is not None and rank == 0) else [None, None]
    ev_render = [torch.cuda.Event() for _ in range(2)]
    ev_gather = [torch.cuda.Event() for _ in range(2)]
    pf = None
    if world > 1 and args.exchange == "peer":
        from j3d_b200.dist import PeerFrames
        try:
            pf = PeerFrames(ctx, H, W, dev, dst=0)
        except RuntimeError as e:  # every rank raises together: CUDA IPC is not available here, gather with NCCL instead
            if rank == 0:
                print(f"bench.py: {e}; using --exchange nccl", file=sys.stderr)
            args.exchange = "nccl"
            comm = torch.cuda.Stream(device=dev)
            gather_lists = [[torch.empty_like(rgba) for _ in range(world)] for _ in range(2)] if rank == 0 else [None, None]

    total = args.warmup + args.steps
    views = frame_views(j, v0, rank, total, world)  # rank r renders frames r, r+N, r+2N ...
    state = {"k": 0}

    def step(v):
        if pf is not None:
            k = pf.begin()
            ctx.render_frame([mesh], [], v, pixels_out=px, rgba_out=pf.target(k))
            pf.end(k)
            return
        k = state["k"]
        state["k"] = k + 1
        b = k & 1
        if comm is not None and k >= 2:
            stream.wait_event(ev_gather[b])  # the gather of frame k - 2 has left this buffer
        ctx.render_frame([mesh], [], v, pixels_out=px, rgba_out=rgba2[b])
        if comm is not None:
            ev_render[b].record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(ev_render[b])
                dist.gather(rgba2[b], gather_lists[b], dst=0)
                ev_gather[b].record(comm)

    def drain():
        if comm is not None:
            stream.wait_stream(comm)

    for v in views[: args.warmup]:
        step(v)
    drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ctx.timings(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for v in views[args.warmup:]:
        step(v)
    drain()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    exchange_ok = True
    if pf is not None:
        # the exchanged frame of the last step equals what this rank rendered (rank 0 checks its own slot and the
        # arrival of everybody else's), and no flag wait ran into its time-out
        exchange_ok = not ctx.stream_wait_timed_out()
        if rank == 0:
            last = pf.frames(pf.k - 1)
            ctx.render_frame([mesh], [], views[-1], pixels_out=px, rgba_out=rgba2[0])
            torch.cuda.synchronize()
            exchange_ok = exchange_ok and bool(torch.equal(last[0], rgba2[0])) and all(int(last[r].abs().sum().item()) != 0 for r in range(world))
    if pf is not None:
        pf.close()
    clocks = sampler.stop() if rank == 0 else None
    tm = ctx.timings(reset=True)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    rays_total = W * H * args.steps * world
    value = rays_total / (ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with HOST buffers (pinned), D2H inside the timed region ----
    # The sweep call a user makes: j3dg_frame_submit / j3dg_frame_wait (pipelined j3dg_render_frame): the
    # device->host copy of frame k (pixel records + RGBA, 74.6 MB) overlaps the kernels of frame k+1.  Every
    # frame's host buffers are complete (frame_wait returned) inside the timed region.
    hpx = [torch.empty((H, W, 32), dtype=torch.uint8).pin_memory() for _ in range(2)]
    hrgba = [torch.empty((H, W), dtype=torch.int32).pin_memory() for _ in range(2)]

    def sweep(vs):
        for k, v in enumerate(vs):
            ctx.frame_submit([mesh], [], v, pixels_out=hpx[k & 1], rgba_out=hrgba[k & 1])
            if k >= 1:
                ctx.frame_wait()
        if vs:
            ctx.frame_wait()

    def timed_sweep():
        sweep(views[: args.warmup])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ctx.readback_bytes(reset=True)
        t0 = time.perf_counter()
        sweep(views[args.warmup:])
        dt = time.perf_counter() - t0
        return dt, ctx.readback_bytes(reset=True) / args.steps

    # (1) every byte of both host buffers copied every frame
    e2e_full_s, full_bytes = timed_sweep()
    # (2) dirty-rectangle readback (j3dg_ctx_set_dirty_rect): the host buffers are persistent per-canvas buffers, so only
    #     the bounding rectangle of the pixels that can differ from what the buffer already holds crosses PCIe; the host
    #     buffers are byte-identical to (1) (tests/test_gpu_parity.py::test_dirty_rect_readback_is_byte_identical and
    #     the check below)
    ref_px, ref_rgba = hpx[(args.steps - 1) & 1].clone(), hrgba[(args.steps - 1) & 1].clone()
    ctx.set_dirty_rect(True)
    e2e_s, d2h_avg = timed_sweep()
    ctx.set_dirty_rect(False)
    dirty_identical = bool(torch.equal(ref_px, hpx[(args.steps - 1) & 1]) and torch.equal(ref_rgba, hrgba[(args.steps - 1) & 1]))
    # the same frames one by one through the synchronous j3dg_render_frame (kernels, then copy)
    for v in views[:3]:  # untimed: the first call allocates the context's own canvas
        ctx.render_frame([mesh], [], v, pixels_out=hpx[0], rgba_out=hrgba[0])
    t0 = time.perf_counter()
    for v in views[args.warmup: args.warmup + min(args.steps, 20)]:
        ctx.render_frame([mesh], [], v, pixels_out=hpx[0], rgba_out=hrgba[0])
    e2e_sync_ms = 1e3 * (time.perf_counter() - t0) / min(args.steps, 20)
    if world > 1:
        t = torch.tensor([e2e_s, e2e_full_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, e2e_full_s = float(t[0].item()), float(t[1].item())
    e2e_value = rays_total / e2e_s / 1e6
    import ctypes
    h2d_bytes = ctypes.sizeof(j.View) + 256  # the view (kernel parameters) + the per-mesh table
    d2h_bytes = int(d2h_avg)

    # ---- interactive-host mode: RGBA-only readback, the pixel records stay in HBM and the host
    # asks for the record under the cursor with j3dg_pick (64 bytes per query) ----
    e2e_rgba = None
    extras = None
    if world == 1:
        def sweep_rgba(vs):
            for k, v in enumerate(vs):
                ctx.frame_submit([mesh], [], v, pixels_out=None, rgba_out=hrgba[k & 1])
                if k >= 1:
                    ctx.frame_wait()
            if vs:
                ctx.frame_wait()
                ctx.pick([mesh], [], vs[-1], np.array([[W // 2, H // 2]], np.int32))
        sweep_rgba(views[: args.warmup])
        t0 = time.perf_counter()
        sweep_rgba(views[args.warmup:])
        dt = time.perf_counter() - t0
        e2e_rgba = {"value": rays_total / dt / 1e6, "unit": "Mrays/s", "ms_per_step": 1e3 * dt / args.steps, "d2h_bytes_per_step": W * H * 4,
                    "api": "j3dg_frame_submit(pixels_out=NULL)/j3dg_frame_wait + j3dg_pick: pixel records stay resident, RGBA only crosses PCIe"}
        # ---- the other query clients of the same BVH, timed through the C ABI ----
        qxy = np.stack(np.meshgrid(np.arange(0, W, 8), np.arange(0, H, 8)), -1).reshape(-1, 2).astype(np.int32)
        ctx.pick([mesh], [], views[-1], qxy)
        t0 = time.perf_counter()
        picks = ctx.pick([mesh], [], views[-1], qxy)
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (cast): algorithmic bytes / live event time ----
    nodes_per_ray, tris_per_ray = ctx.cast_stats([mesh], views[args.warmup])
    bytes_per_ray = nodes_per_ray * info.node_bytes + tris_per_ray * info.triangle_bytes + 32
    cast_ms = tm.cast_ms / max(1, tm.cast_count)
    peak, peak_src = peaks()
    achieved = (W * H * bytes_per_ray) / (cast_ms * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("cast_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "cast stage = cast_kernel (one cooperative launch: lane warps + 8-lane hard-ray groups) + resolve_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray,
                "tris_per_ray": tris_per_ray, "kernel_ms": cast_ms, "kernel_mrays_s": W * H / cast_ms / 1e3}
    cpu = cpu_baseline(j, f, verts, tris, v0) if (world == 1 and not args.no_cpu_baseline) else None
    # the other stages against the same HBM roofline (algorithmic bytes per unit)
    nv = verts.shape[0]
    bvh_bytes = int(info.nr_of_nodes) * info.node_bytes + nt * info.triangle_bytes
    shade_ms = tm.shade_ms / max(1, tm.shade_count)
    stages = {"build": {"ms": build_ms, "algorithmic_bytes": 12 * nt + 12 * nv + bvh_bytes, "mtris_s": nt / build_ms / 1e3 if build_ms else None,
                        "roofline_frac": (12 * nt + 12 * nv + bvh_bytes) / (build_ms * 1e-3) / 1e9 / peak if build_ms else None},
              "shade": {"ms": shade_ms, "algorithmic_bytes": 40 * W * H, "roofline_frac": 40 * W * H / (shade_ms * 1e-3) / 1e9 / peak if shade_ms else None}}

    line = {
        "metric": "primary Mrays/s @1080p (ray cast + shading per frame)", "value": value, "unit": "Mrays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(f, nt), "l2": "inputs_larger_than_l2",
                   "sharding": "replicas only" if world == 1 else f"orbit frames round-robin over {world} ranks, BVH NCCL-broadcast from rank 0, " + ("every rank's shade kernel stores its RGBA into rank 0's HBM over NVLink peer memory (CUDA IPC), stream-ordered arrival/release flags" if args.exchange == "peer" else "RGBA NCCL-gathered on rank 0 every step (second stream, overlapping the kernels of frame k+1)"),
                   "bvh_bytes": int(info.nr_of_nodes) * info.node_bytes + nt * info.triangle_bytes},
        "bvh_build_ms": build_ms, "bvh_nodes": int(info.nr_of_nodes), "frames_per_s": 1e3 * args.steps * world / ms,
        "cast_ms": cast_ms, "shade_ms": tm.shade_ms / max(1, tm.shade_count),
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "j3dg_frame_submit/j3dg_frame_wait (pipelined, pinned persistent host buffers, dirty-rectangle readback: pixel records + RGBA byte-identical to a full copy)",
                "host_buffers_identical_to_full_copy": dirty_identical,
                "full_copy": {"value": rays_total / e2e_full_s / 1e6, "ms_per_step": 1e3 * e2e_full_s / args.steps, "d2h_bytes_per_step": int(full_bytes)},
                "sync_render_frame_ms_per_step": e2e_sync_ms, "mesh_create_ms": e2e_build_ms},
        "gpu_launches": int(tm.kernel_launches),
        "clocks": clocks, "roofline": roofline, "stages": stages,
    }
    if e2e_rgba is not None:
        line["e2e_rgba_only"] = e2e_rgba
    if extras is not None:
        line["queries"] = extras
    if bcast_ms is not None:
        line["bvh_broadcast_ms"] = bcast_ms
        line["exchange"] = {"kind": args.exchange, "verified": exchange_ok, "bytes_per_step_into_rank0": (world - 1) * W * H * 4}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=360)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="B", choices=sorted(WORKLOADS))
    ap.add_argument("--f", type=int, default=0, help="override the icosphere frequency (T = 20 f^2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="N > 1: how every rank's RGBA frame reaches rank 0")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    f = args.f or WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        if args.steps > 12:  # bounded sample: a CPU frame of the 28M mesh takes ~0.1-1 s
            args.steps = 12
        return run_reference(args, f, rank, world)
    from j3d_b200 import build
    if not (ROOT / "j3d_b200" / "libj3dg.so").exists():
        build.build_all()
    return run_b200(args, f, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
