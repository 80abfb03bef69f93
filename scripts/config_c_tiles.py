"""BASELINE configs[2]: ONE frame of a huge mesh rendered by N GPUs — screen bands sharded across ranks
inside the cast / shade kernels (j3dg_ctx_set_screen_shard), BVH built once on rank 0 and NCCL-broadcast,
RGBA bands NCCL-gathered on rank 0 every frame (j3d_b200.dist.gather_bands).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/config_c_tiles.py [--f 3873] [--size 3840x2160] [--frames 8]
  python scripts/config_c_tiles.py --f 1184           # N = 1: no sharding, same code path

Prints one JSON line on rank 0: ms per frame (CUDA events, max over ranks, gather included), Mrays/s (primary +
shadow rays of the whole frame), and whether the gathered frame is bit-identical to the unsharded frame
rendered by rank 0 alone (the parity check of the sharded path at full size)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, torch.distributed as dist
import j3d_b200 as j
from j3d_b200 import dist as jd

ap = argparse.ArgumentParser()
ap.add_argument("--f", type=int, default=3873)
ap.add_argument("--size", default="3840x2160")
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--no-shadow", action="store_true")
ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                help="peer: every rank's shade kernel writes its bands straight into ONE frame in rank 0's HBM (NVLink peer memory); nccl: gather_bands")
args = ap.parse_args()
W, H = (int(t) for t in args.size.split("x"))
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = j.Context(local)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
mc, cav = j.make_matcap(0)
ctx.set_matcap(mc, cav)

nt, nv = 20 * args.f * args.f, 10 * args.f * args.f + 2
t0 = time.time()
build_ms = None
if rank == 0:  # only rank 0 needs the indexed mesh: the others receive the finished BVH (triangle records hold the vertices)
    verts, tris = j.icosphere(args.f)
    gen_s = time.time() - t0
    mesh = ctx.mesh_create(verts, tris)
    ctx.synchronize()
    info = mesh.info()
    build_ms = info.build_ms
    mn, mx = j.compute_bb(verts)
    bb = torch.tensor(np.concatenate([mn, mx]), dtype=torch.float32, device=dev)
    meta = torch.tensor([info.nr_of_nodes], dtype=torch.int64, device=dev)
    del verts, tris
else:
    bb = torch.zeros(6, dtype=torch.float32, device=dev)
    meta = torch.zeros(1, dtype=torch.int64, device=dev)
bcast_ms = 0.0
if world > 1:
    dist.broadcast(meta, 0); dist.broadcast(bb, 0)
    if rank != 0:
        mesh = ctx.mesh_create_empty(nv, nt, int(meta.item()))
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); jd.broadcast_bvh(mesh, src=0, device=dev); e1.record(); torch.cuda.synchronize()
    bcast_ms = e0.elapsed_time(e1)
bbh = bb.cpu().numpy()
flags = j.DEFAULT_FLAGS | (0 if args.no_shadow else j.SHADOW)
v0 = j.make_view(W, H, bbh[:3], bbh[3:], flags)
views = [j.orbit_view(v0, 25.0 * k) for k in range(args.warmup + args.frames)]
px = torch.zeros((H, W, 32), dtype=torch.uint8, device=dev)
rgba = torch.zeros((H, W), dtype=torch.int32, device=dev)

pf = jd.PeerFrames(ctx, H, W, dev, dst=0, shared_frame=True) if (world > 1 and args.exchange == "peer") else None

def frame(v):
    if pf is not None:
        k = pf.begin()
        ctx.render_frame([mesh], [], v, pixels_out=px, rgba_out=pf.target(k))
        pf.end(k)
        return pf.frames(k)[0] if rank == 0 else None
    ctx.render_frame([mesh], [], v, pixels_out=px, rgba_out=rgba)
    return jd.gather_bands(rgba, dst=0) if world > 1 else rgba

ctx.set_screen_shard(rank, world)
for v in views[: args.warmup]:
    frame(v)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ctx.timings(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for v in views[args.warmup:]:
    whole = frame(v)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
tm = ctx.timings(reset=True)
rays = torch.tensor([float(tm.rays)], dtype=torch.float64, device=dev)
stage = torch.tensor([tm.cast_ms, tm.shade_ms], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(rays, op=dist.ReduceOp.SUM); dist.all_reduce(stage, op=dist.ReduceOp.MAX)
# parity of the sharded path: rank 0 renders the last frame alone, unsharded
identical = None
if rank == 0:
    torch.cuda.synchronize()
    got = whole.clone()
    ctx.set_screen_shard(0, 1)
    ctx.render_frame([mesh], [], views[-1], pixels_out=px, rgba_out=rgba)
    torch.cuda.synchronize()
    identical = bool(torch.equal(got, rgba))
    p = px.cpu().numpy().view(j.PIXEL_DTYPE).reshape(H, W)
    hit = p["object_id"] != 0xFFFFFFFF
    print(json.dumps({"config": f"icosphere f={args.f} ({nt} triangles), {W}x{H}, shadows {'off' if args.no_shadow else 'on'}, screen bands of 32 rows over {world} GPU(s)",
                      "n_gpus": world, "frames": args.frames, "ms_per_frame": ms.item() / args.frames, "frames_per_s": 1e3 * args.frames / ms.item(),
                      "mrays_per_s": rays.item() / ms.item() / 1e3, "rays_per_frame": rays.item() / args.frames,
                      "cast_ms_max_rank": stage[0].item() / args.frames, "shade_ms_max_rank": stage[1].item() / args.frames,
                      "bvh_build_ms": build_ms, "bvh_broadcast_ms": bcast_ms, "bvh_bytes": int(info.nr_of_nodes) * info.node_bytes + nt * info.triangle_bytes,
                      "gathered_equals_unsharded": identical, "hit_pixels": int(hit.sum()), "shadowed": int((p["mark"][hit] & 1).sum()),
                      "mesh_generate_s": gen_s, "exchange": args.exchange if world > 1 else None,
                      "exchange_timed_out": bool(ctx.stream_wait_timed_out()) if pf is not None else None}))
if pf is not None:
    pf.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
