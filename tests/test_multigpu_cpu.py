"""CPU tests of the N>1 host logic: slab partition maths, and the broadcast / gather plumbing over torch.distributed
with the gloo backend at world_size 2 (the B200 box runs the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest


def test_slab_ranges_tile_the_volume():
    from rrpy import multigpu as M
    for Z in (1, 7, 128, 512, 1000):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            sizes = []
            for r in range(world):
                z0, z1 = M.slab_range(r, world, Z)
                assert z0 == prev and z1 >= z0
                sizes.append(z1 - z0)
                prev = z1
            assert prev == Z and max(sizes) - min(sizes) <= 1
    assert M.halo(0.01, 512) == 8 and M.halo(0.01, 100) == 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from rrpy import multigpu as M
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # frame broadcast: rank 0 holds the frame set, everybody ends up with it
        g = torch.Generator().manual_seed(5)
        color_src = torch.randint(0, 255, (2, 6, 8, 3), dtype=torch.uint8, generator=g)
        depth_src = torch.rand((2, 5, 7), generator=g)
        color = color_src.clone() if rank == 0 else torch.zeros_like(color_src)
        depth = depth_src.clone() if rank == 0 else torch.zeros_like(depth_src)
        M.broadcast_frames(dist, color, depth, 0)
        ok = bool(torch.equal(color, color_src) and torch.equal(depth, depth_src))
        # the packed, double-buffered broadcaster of the slab path: two frame sets in flight, consumed in order
        fb = M.FrameBroadcaster(dist, "cpu", color_src.numel(), depth_src.numel() * 4, src=0)
        sets = [(color_src, depth_src), (255 - color_src, depth_src * 2.0)]
        for c_, d_ in sets:
            fb.issue(c_ if rank == 0 else None, d_ if rank == 0 else None)
        for c_, d_ in sets:
            packed, slot = fb.consume()
            gc, gd = fb.unpack(packed, tuple(color_src.shape), tuple(depth_src.shape))
            ok = ok and bool(torch.equal(gc, c_) and torch.equal(gd, d_))
            fb.release(slot)
        ok = ok and fb.in_flight() == 0
        # record gather: each rank "hits" a different subset of 6 pixels at rank-dependent step indices
        n = 6
        rec = np.zeros((n, M.RECORD_FLOATS), np.float32)
        steps = np.full(n, 0xFFFFFFFF, np.uint32)
        mine = [0, 2, 4] if rank == 0 else [2, 3, 4]
        for px in mine:
            steps[px] = 10 + 5 * rank + (0 if px != 4 else 20 * (1 - rank))      # pixel 4: rank 1 is earlier, pixel 2: rank 0
        rec[:, 5] = steps.view(np.float32)
        rec[:, 0] = rank + 1                                                      # colour tags the owner
        out = M.gather_records(dist, torch.from_numpy(rec), dst=0)
        if rank == 0:
            parts = out.numpy()
            comp, win = M.composite_reference(parts)
            q.put((ok, parts.shape, win.tolist(), comp[:, 0].tolist()))
        else:
            q.put((ok, None, None, None))
    finally:
        dist.destroy_process_group()


def test_gloo_broadcast_and_gather_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] for r in results), "broadcast did not deliver the frame set"
    root = [r for r in results if r[1] is not None][0]
    assert root[1] == (2, 6, 8)
    # pixel 0: only rank 0; 1: nobody (first = rank 0); 2: rank 0 earlier; 3: only rank 1; 4: rank 1 earlier; 5: nobody
    assert root[2] == [0, 0, 0, 1, 1, 0]
    assert root[3] == [1.0, 1.0, 1.0, 2.0, 2.0, 1.0]


def test_balanced_slabs_tile_the_volume_and_follow_the_occupancy():
    from rrpy import multigpu as M
    Z, plane = 64, 64 * 64
    # occupied bricks only in z in [24, 40): the middle of the volume
    ranges = np.array([[0, 16, 0, 16, z, z + 8] for z in range(0, 64, 8)], np.int32)
    occupied = np.array([3, 4], np.uint32)
    for world in (1, 2, 3, 4, 8):
        slabs = M.balanced_slabs(world, Z, plane, ranges, occupied)
        assert slabs[0][0] == 0 and slabs[-1][1] == Z and len(slabs) == world
        assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:])) and all(z1 > z0 for z0, z1 in slabs)
    four = M.balanced_slabs(4, Z, plane, ranges, occupied, compute_to_fill=45.0)
    thick = [z1 - z0 for z0, z1 in four]
    assert thick[0] > thick[1] and thick[3] > thick[2], f"edge slabs should be thicker than the occupied middle: {four}"
    # no occupancy -> equal thickness, like slab_range
    assert M.balanced_slabs(4, Z, plane, ranges, np.zeros(0, np.uint32)) == [M.slab_range(r, 4, Z) for r in range(4)]
    # world == Z: one slice each
    assert M.balanced_slabs(8, 8, 4, ranges[:1], np.array([0], np.uint32)) == [(i, i + 1) for i in range(8)]
