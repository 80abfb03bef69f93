"""N>1 host logic on CPU: world_size-2 gloo processes exercising the partition + gather helpers
the multi-GPU bench uses (frames round-robin, row bands, tiles, gather to rank 0, packed-word max)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from j3d_b200 import dist as jd


def test_partitions_cover_exactly_once():
    for world in (1, 2, 3, 8):
        frames = sorted(sum((jd.frames_for_rank(360, r, world) for r in range(world)), []))
        assert frames == list(range(360))
        bands = jd.row_bands(1080, world)
        assert bands[0][0] == 0 and bands[-1][1] == 1079 and all(bands[i][1] + 1 == bands[i + 1][0] for i in range(world - 1))
        rows = sorted(y for r in range(world) for (a, b) in jd.bands_for_rank(1080, r, world) for y in range(a, b + 1))
        assert rows == list(range(1080))
        cover = np.zeros((270, 480), np.int32)
        for r in range(world):
            for (x0, y0, x1, y1) in jd.tiles_for_rank(480, 270, r, world, tile=64):
                cover[y0:y1 + 1, x0:x1 + 1] += 1
        assert (cover == 1).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h, w = 37, 16
        bands = jd.row_bands(h, world)
        full = torch.arange(h * w, dtype=torch.int32).reshape(h, w)
        y0, y1 = bands[rank]
        got = jd.gather_rows(full[y0:y1 + 1].clone(), bands, dst=0)
        frames = jd.gather_frames(torch.full((4, 4), rank, dtype=torch.int32), dst=0)
        packed = torch.tensor([(10 + rank) << 32 | 5, (20 - rank) << 32 | rank], dtype=torch.int64)
        jd.allreduce_max_u64(packed)
        # screen sharding of one frame: every rank holds a frame with only its own 32-row bands valid
        for hh in (37, 64, 200):
            truth = torch.arange(hh * w, dtype=torch.int32).reshape(hh, w)
            mine = torch.full_like(truth, -1)
            for (b0, b1) in jd.bands_for_rank(hh, rank, world):
                mine[b0:b1 + 1] = truth[b0:b1 + 1]
            whole = jd.gather_bands(mine, dst=0)
            if rank == 0:
                assert torch.equal(whole, truth)
        if rank == 0:
            assert torch.equal(got, full)
            assert [int(f[0, 0]) for f in frames] == list(range(world))
        assert packed.tolist() == [(10 + world - 1) << 32 | 5, 20 << 32 | 0]
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    assert sorted(out.keys()) == [0, 1]


def test_peer_exchange_layout():
    """Slot / flag arithmetic of the NVLink peer-memory frame exchange (dist.PeerFrames): frames never overlap,
    both slots of every rank are distinct, the flag words lie behind the frames."""
    from j3d_b200.dist import peer_flags_offset, peer_slot_offset
    fb = 1920 * 1080 * 4
    for world in (1, 2, 8):
        offs = sorted(peer_slot_offset(s, r, world, fb) for s in (0, 1) for r in range(world))
        assert offs == [i * fb for i in range(2 * world)]
        fo = peer_flags_offset(world, fb)
        assert fo >= 2 * world * fb and fo % 256 == 0


def test_reference_arm_under_torchrun():
    """bench.py --impl reference launched like the driver launches it for N > 1: rank 0 alone runs the reference's CPU
    renderer and prints ONE JSON line, the other rank exits 0 without work (config A here so that it takes seconds)."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    import pytest
    root = Path(__file__).resolve().parents[1]
    if not (root / "oracle" / "_ref" / "libj3d_ref.so").exists():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           str(root / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--workload", "A"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(root))
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


class _RecordingCtx:
    """Stands in for j3d_b200.Context: records the stream-ordered flag operations PeerFrames enqueues (no GPU)."""

    def __init__(self, rank, fail_open=False):
        self.rank, self.fail_open, self.ops = rank, fail_open, []

    def peer_alloc(self, nbytes):
        self.nbytes = nbytes
        return 0x10000000, b"H" * 64

    def peer_open(self, handle):
        assert handle == b"H" * 64
        if self.fail_open:
            raise RuntimeError("cudaIpcOpenMemHandle failed")
        return 0x20000000

    def peer_close(self, ptr):
        self.ops.append(("close", ptr))

    def peer_free(self, ptr):
        self.ops.append(("free", ptr))

    def stream_signal(self, ptr, value):
        self.ops.append(("signal", ptr, value))

    def stream_wait_geq(self, ptr, n, value):
        self.ops.append(("wait", ptr, n, value))

    def synchronize(self):
        pass


def _peer_worker(rank, world, port, out, fail):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h, w = 8, 16
        fb = h * w * 4
        ctx = _RecordingCtx(rank, fail_open=fail and rank == 1)
        if fail:
            with pytest.raises(RuntimeError):  # every rank raises together: nobody is left waiting in a collective
                jd.PeerFramesPy(ctx, h, w, torch.device("cpu"), dst=0)
            out[rank] = [op[0] for op in ctx.ops]
            return
        pf = jd.PeerFramesPy(ctx, h, w, torch.device("cpu"), dst=0)
        base = 0x10000000 if rank == 0 else 0x20000000
        flags = base + jd.peer_flags_offset(world, fb)
        targets = []
        for step in range(4):
            k = pf.begin()
            targets.append(pf.target(k) - base)
            pf.arrive(k)
            ctx.ops.append(("consume", k))   # what rank 0 enqueues between arrival and release (a no-op marker elsewhere)
            pf.release(k)
        assert targets == [jd.peer_slot_offset(k & 1, rank, world, fb) for k in range(4)]
        want = []
        for k in range(4):
            s = k & 1                                                        # flag words per slot: arrived[slot][rank], then released[slot]
            if k >= 2:
                want.append(("wait", flags + 4 * (2 * world + s), 1, k - 1))   # slot k & 1 released by rank 0 (frame k - 2 consumed)
            want.append(("signal", flags + 4 * (s * world + rank), k + 1))     # my frame k has arrived
            if rank == 0:
                want.append(("wait", flags + 4 * s * world, world, k + 1))      # rank 0: every rank's frame k
            want.append(("consume", k))
            if rank == 0:
                want.append(("signal", flags + 4 * (2 * world + s), k + 1))    # ... consumed: release
        assert ctx.ops == want, (ctx.ops, want)
        # two frames in flight: the frames of slot 1 are handed over on a second context's stream
        lane0, lane1 = _RecordingCtx(rank), _RecordingCtx(rank)
        two = jd.PeerFramesPy(lane0, h, w, torch.device("cpu"), dst=0)
        two.set_lane(1, lane1)
        for step in range(4):
            k = two.begin()
            two.arrive(k)
            two.release(k)
        assert all(op[1] in (flags + 4 * rank, flags, flags + 8 * world) for op in lane0.ops if op[0] in ("signal", "wait"))
        assert all(op[1] in (flags + 4 * (world + rank), flags + 4 * world, flags + 4 * (2 * world + 1)) for op in lane1.ops if op[0] in ("signal", "wait"))
        assert len(lane0.ops) == len(lane1.ops) > 0
        # three slots, three lanes: frame k on lane k mod 3, begin(k) waits for the release of frame k - 3
        lanes = [_RecordingCtx(rank) for _ in range(3)]
        three = jd.PeerFramesPy(lanes[0], h, w, torch.device("cpu"), dst=0, nslots=3)
        flags3 = base + jd.peer_flags_offset(world, fb, 3)
        for s in (1, 2):
            three.set_lane(s, lanes[s])
        for step in range(7):
            k = three.begin()
            assert three.target(k) - base == jd.peer_slot_offset(k % 3, rank, world, fb)
            three.arrive(k)
            three.release(k)
        for s in range(3):
            want3 = []
            for k in range(s, 7, 3):
                if k >= 3:
                    want3.append(("wait", flags3 + 4 * (3 * world + s), 1, k - 2))
                want3.append(("signal", flags3 + 4 * (s * world + rank), k + 1))
                if rank == 0:
                    want3.append(("wait", flags3 + 4 * s * world, world, k + 1))
                    want3.append(("signal", flags3 + 4 * (3 * world + s), k + 1))
            assert lanes[s].ops == want3, (s, lanes[s].ops, want3)
        shared = jd.PeerFramesPy(_RecordingCtx(rank), h, w, torch.device("cpu"), dst=0, shared_frame=True)
        assert shared.target(0) - base == 0 and shared.target(1) - base == jd.peer_slot_offset(1, 0, world, fb)  # one frame per slot
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("fail", [False, True])
def test_peer_frames_protocol_world2_gloo(fail):
    """The hand-over protocol of the NVLink peer-memory frame exchange (dist.PeerFramesPy, the executable specification of
    j3dg_frames_* in csrc/group.cu) with two ranks and a recording
    context: slots, arrival / release flag values and their order; and a failing CUDA-IPC open on one rank makes BOTH
    ranks raise (so the caller can fall back to an NCCL gather) after cleaning up."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, out, fail)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    assert sorted(out.keys()) == [0, 1]
    if fail:
        assert out[0] == ["free"] and out[1] == []   # rank 0 releases its buffer, rank 1 never mapped it
