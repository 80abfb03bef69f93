"""Python side of the multi-GPU path: one process per GPU.  The plumbing itself — NCCL communicator, mesh / BVH
broadcast, the peer-memory frame exchange — lives behind the C ABI (include/j3dg.h "multi-GPU", csrc/group.cu, driven
from plain C++ by tests/cpp/group_ranks.cpp); this module is a thin binding for torchrun-launched Python processes
(torch.distributed only ships the 128-byte NCCL id and serves the gloo CPU tests) plus the pure partition functions.
The path shards only where it does so naturally: the mesh is replicated, the BVH is built once and
broadcast, orbit frames or screen bands are partitioned, rank 0 ends up with the results.  There is no collective
inside the ray cast itself.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import capi


def make_group(ctx) -> "capi.Group":
    """j3dg_group for this torchrun-launched process: rank 0 creates the NCCL id, torch.distributed ships it."""
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [capi.group_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return capi.Group(ctx, rank, world, box[0])


# ---- partitioning (pure functions; tested on CPU with gloo, world_size 2) -------------------
def frames_for_rank(n_frames: int, rank: int, world: int) -> list[int]:
    """Orbit sweep: frame k goes to rank k mod world (round robin keeps neighbouring poses apart,
    so every rank sees the same mix of cheap and expensive views)."""
    return list(range(rank, n_frames, world))


def tiles(width: int, height: int, tile: int = 64) -> list[tuple[int, int, int, int]]:
    """Inclusive (x0, y0, x1, y1) rectangles covering the canvas, row-major."""
    out = []
    for y in range(0, height, tile):
        for x in range(0, width, tile):
            out.append((x, y, min(x + tile, width) - 1, min(y + tile, height) - 1))
    return out


def _morton2(x: int, y: int) -> int:
    r = 0
    for b in range(16):
        r |= ((x >> b) & 1) << (2 * b) | ((y >> b) & 1) << (2 * b + 1)
    return r


def tiles_for_rank(width: int, height: int, rank: int, world: int, tile: int = 64):
    """Screen tiles interleaved over ranks in Morton order (silhouette-heavy regions are spread)."""
    ts = sorted(tiles(width, height, tile), key=lambda t: _morton2(t[0] // tile, t[1] // tile))
    return ts[rank::world]


def row_bands(height: int, world: int) -> list[tuple[int, int]]:
    """Contiguous inclusive row bands [y0, y1], as equal as possible."""
    base, rem = divmod(height, world)
    out, y = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((y, y + n - 1))
        y += n
    return out


# ---- collectives -----------------------------------------------------------------------------
class _DevBuf:
    """Zero-copy view of a raw device allocation for torch (CUDA array interface)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


def device_bytes(ptr: int, nbytes: int, device) -> torch.Tensor:
    return torch.as_tensor(_DevBuf(ptr, nbytes), device=device)


def broadcast_bvh(mesh, src: int = 0, device=None, chunk: int = 1 << 30):
    """Broadcast the built BVH (wide nodes + triangle records) in place from `src` with torch.distributed — kept for
    the NCCL-gather comparison arm; the product path is Group.broadcast_mesh (j3dg_group_broadcast_mesh)."""
    for kind in (0, 1):
        ptr, nbytes = mesh.bvh_buffer(kind)
        if not nbytes:
            continue
        t = device_bytes(ptr, nbytes, device)
        for off in range(0, nbytes, chunk):
            dist.broadcast(t[off: off + chunk], src)


def gather_rows(local: torch.Tensor, bands: list[tuple[int, int]], dst: int = 0) -> torch.Tensor | None:
    """Gather contiguous row bands of an image (rank r owns rows bands[r]) on `dst`."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == dst:
        parts = [torch.empty((b[1] - b[0] + 1,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) for b in bands]
        parts[dst] = local
        reqs = [dist.irecv(parts[r], src=r) for r in range(world) if r != dst]
        for q in reqs:
            q.wait()
        return torch.cat(parts, dim=0)
    dist.send(local.contiguous(), dst=dst)
    return None


def gather_frames(frame: torch.Tensor, dst: int = 0) -> list[torch.Tensor] | None:
    """One frame per rank -> list of frames on `dst` (orbit sweep step)."""
    world = dist.get_world_size()
    out = [torch.empty_like(frame) for _ in range(world)] if dist.get_rank() == dst else None
    dist.gather(frame, out, dst=dst)
    return out


def allreduce_max_u64(t: torch.Tensor) -> torch.Tensor:
    """Per-pixel max of the packed (depth, id) splat words for a point-range-sharded cloud.
    NCCL has no uint64 max; the words are < 2^63 (positive float bits in the top half), so the
    signed int64 max is identical."""
    dist.all_reduce(t.view(torch.int64), op=dist.ReduceOp.MAX)
    return t


# ---- screen sharding of ONE frame (BASELINE configs[2]; j3dg_ctx_set_screen_shard) -----------
BAND_ROWS = 32  # = J3DG_SHARD_BAND_ROWS


def bands_for_rank(height: int, rank: int, world: int) -> list[tuple[int, int]]:
    """Inclusive row ranges [y0, y1] of the bands rank owns: band b belongs to rank b mod world."""
    nb = (height + BAND_ROWS - 1) // BAND_ROWS
    return [(b * BAND_ROWS, min((b + 1) * BAND_ROWS, height) - 1) for b in range(rank, nb, world)]


def pack_bands(img: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's bands of `img` ([H, ...]) as one [per, BAND_ROWS, ...] tensor (per = bands per rank,
    rounded up; missing rows are zero)."""
    h = img.shape[0]
    nb = (h + BAND_ROWS - 1) // BAND_ROWS
    per = (nb + world - 1) // world
    if h == nb * BAND_ROWS and nb == per * world:
        return img.view((per, world, BAND_ROWS) + tuple(img.shape[1:]))[:, rank].contiguous()
    out = img.new_zeros((per, BAND_ROWS) + tuple(img.shape[1:]))
    for l, (y0, y1) in enumerate(bands_for_rank(h, rank, world)):
        out[l, : y1 - y0 + 1] = img[y0: y1 + 1]
    return out


def gather_bands(img: torch.Tensor, dst: int = 0) -> torch.Tensor | None:
    """Every rank holds a frame whose own bands are valid (a frame rendered with set_screen_shard(rank, world));
    returns the complete frame on `dst`.  One gather of 1/world of the frame per rank; on `dst` the gathered
    [world, per, 32, ...] block is the frame with the two leading axes swapped."""
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = pack_bands(img, rank, world)
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == dst else None
    dist.gather(mine, parts, dst=dst)
    if rank != dst:
        return None
    g = torch.stack(parts, dim=0)  # [world, per, 32, ...]
    g = g.transpose(0, 1).reshape((-1,) + tuple(img.shape[1:]))
    return g[: img.shape[0]]


# ---- frames handed to rank 0 through NVLink peer memory instead of a collective (include/j3dg.h, csrc/group.cu, csrc/peer.cu) ----
class PeerFrames:
    """Every rank renders its frame STRAIGHT INTO rank `dst`'s HBM (the shade kernel's stores travel over NVLink);
    no gather kernel competes with the persistent cast kernel for SMs.  A thin wrapper over j3dg_frames_* (the
    protocol is in include/j3dg.h).  Double-buffered, all stream-ordered:

        k = pf.begin()                 # stream waits until dst has RELEASED the frame that lived in slot k & 1
        ctx.render_frame(..., rgba_out=pf.target(k))
        pf.arrive(k)                   # signal arrival; on dst the stream then waits for every rank's frame k
        ... dst enqueues its consumer of pf.frames(k) on the same stream (copy to host, encode, compare) ...
        pf.release(k)                  # dst: everything enqueued so far has read the slot; peers may overwrite it

    shared_frame = True: ONE frame per slot that all ranks write disjoint rows of (j3dg_ctx_set_screen_shard)."""

    def __init__(self, ctx, height: int, width: int, device, dst: int = 0, shared_frame: bool = False, group=None, nslots: int = 2):
        self.ctx, self.h, self.w, self.device, self.dst, self.shared, self.nslots = ctx, height, width, device, dst, shared_frame, nslots
        self._own_group = group is None
        self.group = make_group(ctx) if group is None else group
        self.rank, self.world = self.group.rank, self.group.world
        self.frame_bytes = height * width * 4
        try:
            self.f = self.group.frames(width, height, dst, shared_frame, nslots)
        except capi.J3dgError as e:  # raised on every rank together (the set-up is collective)
            if self._own_group:
                self.group.destroy()
            raise RuntimeError(str(e)) from e
        self.k = 0

    def set_lane(self, slot: int, ctx):
        """Frames in flight: the frames of this slot (k mod nslots) are rendered by this context (j3dg_frames_set_lane)."""
        self.f.set_lane(slot, ctx)

    def begin(self) -> int:
        return self.f.begin()

    def target(self, k: int) -> int:
        return self.f.target(k)

    def arrive(self, k: int):
        self.f.arrive(k)
        self.k = k + 1

    def release(self, k: int):
        """dst only (a no-op elsewhere): the consumer work of frame k is enqueued; the slot may be overwritten."""
        self.f.release(k)

    def end(self, k: int):
        self.arrive(k)
        self.release(k)

    def frames(self, k: int) -> torch.Tensor:
        assert self.rank == self.dst
        n = 1 if self.shared else self.world
        t = device_bytes(self.f.view(k), n * self.frame_bytes, self.device)
        return t.view(torch.int32).view(n, self.h, self.w)

    def close(self):
        self.f.destroy()
        if self._own_group:
            self.group.destroy()


def peer_slot_offset(slot: int, rank: int, world: int, frame_bytes: int) -> int:
    """Byte offset of (slot, rank)'s frame inside the exchange buffer: [nslots][world][frame]."""
    return (slot * world + rank) * frame_bytes


def peer_flags_offset(world: int, frame_bytes: int, nslots: int = 2) -> int:
    """The flag words follow the frames (256-byte aligned): arrived[slot][rank], then released[slot]."""
    return (nslots * world * frame_bytes + 255) & ~255


class PeerFramesPy:
    """Executable specification of the hand-over protocol that csrc/group.cu implements (j3dg_frames_*), written on
    the flag primitives of the ABI (j3dg_peer_alloc / j3dg_peer_open / j3dg_stream_signal / j3dg_stream_wait_geq) and
    torch.distributed.  tests/test_dist_gloo.py runs it with two gloo ranks and a recording context to pin the slots,
    the flag values and their order without a GPU; the product path is PeerFrames above."""

    def __init__(self, ctx, height: int, width: int, device, dst: int = 0, shared_frame: bool = False, nslots: int = 2):
        """nslots: frames that may be in flight (frame k lives in slot k mod nslots).
        shared_frame = False: one frame per rank and slot (orbit sweep: every rank renders its own frame).
        shared_frame = True : ONE frame per slot that all ranks write disjoint rows of (screen sharding,
        j3dg_ctx_set_screen_shard: each rank's shade kernel writes only its own bands) — the gather disappears."""
        self.ctx, self.h, self.w, self.device, self.dst = ctx, height, width, device, dst
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.shared = shared_frame
        self.frame_bytes = height * width * 4
        self.nslots = nslots
        self.flags_off = peer_flags_offset(self.world, self.frame_bytes, nslots)
        self.nbytes = self.flags_off + 4 * nslots * (self.world + 1) + 256
        # set-up is collective and must not leave a rank behind: every rank reports whether its step worked, the group
        # agrees (MIN), and on any failure everybody raises the same error (callers may then fall back to an NCCL gather)
        handle = [None]
        ok, err = 1, ""
        self.base = 0
        if self.rank == dst:
            try:
                self.base, h = ctx.peer_alloc(self.nbytes)
                handle[0] = h
            except Exception as e:  # noqa: BLE001
                ok, err = 0, str(e)
        dist.broadcast_object_list(handle, src=dst)
        if self.rank != dst and handle[0] is not None:
            try:
                self.base = ctx.peer_open(handle[0])
            except Exception as e:  # noqa: BLE001
                ok, err = 0, str(e)
        elif handle[0] is None:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if self.base:
                try:
                    (ctx.peer_free if self.rank == dst else ctx.peer_close)(self.base)
                except Exception:  # noqa: BLE001
                    pass
                self.base = 0
            raise RuntimeError("peer-memory frame exchange unavailable on at least one rank" + (f": {err}" if err else ""))
        self.k = 0
        self.lane = [ctx] * nslots
        dist.barrier()

    # flag words: arrived[slot][rank] (nslots * world words), then released[slot] (nslots words) — one set per slot, so
    # that the slots can be driven from different streams (frames in flight) and every word only ever grows
    def _arrived(self, slot: int, r: int) -> int:
        return self.base + self.flags_off + 4 * (slot * self.world + r)

    def _released(self, slot: int) -> int:
        return self.base + self.flags_off + 4 * (self.nslots * self.world + slot)

    def set_lane(self, slot: int, ctx):
        """The frames of this slot (k mod nslots) are rendered and handed over on this context's stream."""
        self.lane[slot] = ctx

    def begin(self) -> int:
        k = self.k
        s = k % self.nslots
        if k >= self.nslots:  # frame k - nslots lived in this slot: dst must have released it (released[slot] = last consumed frame + 1)
            self.lane[s].stream_wait_geq(self._released(s), 1, k - self.nslots + 1)
        return k

    def target(self, k: int) -> int:
        return self.base + peer_slot_offset(k % self.nslots, 0 if self.shared else self.rank, self.world, self.frame_bytes)

    def arrive(self, k: int):
        s = k % self.nslots
        ctx = self.lane[s]
        ctx.stream_signal(self._arrived(s, self.rank), k + 1)
        if self.rank == self.dst:
            ctx.stream_wait_geq(self._arrived(s, 0), self.world, k + 1)
        self.k = k + 1

    def release(self, k: int):
        """dst only (a no-op elsewhere): the consumer work of frame k is enqueued; the slot may be overwritten."""
        if self.rank == self.dst:
            self.lane[k % self.nslots].stream_signal(self._released(k % self.nslots), k + 1)

    def end(self, k: int):
        self.arrive(k)
        self.release(k)

    def frames(self, k: int) -> torch.Tensor:
        assert self.rank == self.dst
        off = peer_slot_offset(k % self.nslots, 0, self.world, self.frame_bytes)
        n = 1 if self.shared else self.world
        t = device_bytes(self.base + off, n * self.frame_bytes, self.device)
        return t.view(torch.int32).view(n, self.h, self.w)

    def close(self):
        self.ctx.synchronize()
        dist.barrier()
        if self.rank == self.dst:
            dist.barrier()  # the others unmap first
            self.ctx.peer_free(self.base)
        else:
            self.ctx.peer_close(self.base)
            dist.barrier()
        self.base = 0
