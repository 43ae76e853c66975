"""Host-side plumbing of the z-slab multi-GPU path: one process per GPU, torch.distributed for the two
exchange steps the path has (frame broadcast before pre-processing, one gather of partial ray records per view).
Backend-agnostic on purpose: the CPU test suite drives it with gloo and CPU tensors, the B200 box with NCCL over
NVLink. No computation happens here — slabs are integrated / marched / composited by librr_b200.so."""
from __future__ import annotations

import numpy as np

RECORD_FLOATS = 8          # RR_PARTIAL_RECORD_BYTES / 4


def slab_range(rank: int, world: int, Z: int):
    """Contiguous z-slices [z0, z1) of `rank`; the remainder is spread over the first ranks. Slabs tile [0, Z)."""
    base, rem = divmod(int(Z), int(world))
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def balanced_slabs(world: int, Z: int, plane_voxels: int, brick_ranges, occupied, compute_to_fill: float = 45.0, halo_slices: int = 0):
    """Contiguous z-slabs of (nearly) equal integrate cost instead of equal thickness: occupied bricks cluster around the
    captured subject, so equal slabs leave the outer ranks idle while the middle ones work (measured on 4 B200: 0.05 ms
    on the edge slab against 0.11 ms in the middle). Cost of slice z = plane_voxels (the clear stream) +
    compute_to_fill * (voxels of occupied bricks in the slice); compute_to_fill is the measured per-voxel cost ratio of
    the brick evaluation against the clear. A slab owner also integrates `halo_slices` slices on either side of
    its slab (rr_integrate), which weighs heavily on thin slabs: the cost of a slab is the cost of slab + halo, and the
    boundaries minimise the largest one (bisection on the bound, greedy fill). Deterministic in its inputs, and every rank
    holds the same brick counters (pre-processing is replicated), so all ranks derive the same boundaries without
    communication. Returns [(z0, z1)] * world tiling [0, Z), every slab non-empty."""
    Z, world, h = int(Z), int(world), max(0, int(halo_slices))
    assert 1 <= world <= Z
    cost = np.full(Z, float(plane_voxels), np.float64)
    rr = np.asarray(brick_ranges, np.int64)[np.asarray(occupied, np.int64)] if len(occupied) else np.zeros((0, 6), np.int64)
    for x0, x1, y0, y1, z0, z1 in rr:
        cost[max(0, z0):min(Z, z1)] += compute_to_fill * float((x1 - x0) * (y1 - y0))
    cum = np.concatenate([[0.0], np.cumsum(cost)])

    def slab_cost(z0, z1):
        return cum[min(Z, z1 + h)] - cum[max(0, z0 - h)]

    def fill(bound):
        """Greedy: every slab as thick as the bound allows (at least one slice); returns the boundaries or None if more
        than `world` slabs would be needed."""
        b = [0]
        while b[-1] < Z:
            if len(b) > world:
                return None
            z0 = b[-1]
            z1 = z0 + 1
            while z1 < Z and slab_cost(z0, z1 + 1) <= bound:
                z1 += 1
            b.append(z1)
        return b

    lo, hi = 0.0, float(cum[-1])
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if fill(mid) is None:
            lo = mid
        else:
            hi = mid
    bounds = fill(hi)
    # fewer slabs than ranks (tiny volumes): split the thickest slabs until every rank has one
    while len(bounds) - 1 < world:
        k = int(np.argmax(np.diff(bounds)))
        assert bounds[k + 1] - bounds[k] >= 2
        bounds.insert(k + 1, (bounds[k] + bounds[k + 1]) // 2)
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world)]


def halo(limit: float, Z: int) -> int:
    """Slices a slab owner recomputes past its slab on each side (rr_integrate does this internally)."""
    return int(np.ceil(np.float32(limit) * np.float32(Z))) + 2


def broadcast_frames(dist, color, depth, src=0):
    """One frame set (colour uint8 [N][CH][CW][3], depth float32 [N][H][W]) from the ingest rank to every rank.
    Tensors must be preallocated on every rank; returns them for convenience."""
    dist.broadcast(color, src)
    dist.broadcast(depth, src)
    return color, depth


class FrameBroadcaster:
    """The per-frame-set collective of the slab path: ONE broadcast of the packed raw frame set
    (colour bytes then depth bytes, like one server message) from the ingest rank, double-buffered so that the broadcast
    of frame set i+1 runs while frame set i is pre-processed and integrated.

      issue(color, depth)    enqueue the broadcast of the next frame set (sources are read on the ingest rank only:
                             device tensors, or pinned host tensors for the end-to-end path)
      consume(stream)        make `stream` wait for the oldest broadcast in flight; returns (packed uint8 tensor, slot)
      release(slot, stream)  the consumer's last read of the slot has been enqueued on `stream`

    On CPU tensors (gloo tests) the stream/event arguments are ignored."""

    def __init__(self, dist, device, color_bytes, depth_bytes, src=0):
        import torch
        assert color_bytes % 4 == 0, "depth must stay 4-byte aligned behind the colour bytes"
        self.dist, self.src, self.cb, self.db = dist, src, int(color_bytes), int(depth_bytes)
        self.cuda = torch.device(device).type == "cuda"
        self.device = device
        self.rank = dist.get_rank()
        self.buf = [torch.empty(self.cb + self.db, dtype=torch.uint8, device=device) for _ in range(2)]
        self.buf_color = [t[:self.cb] for t in self.buf]
        self.buf_depth = [t[self.cb:] for t in self.buf]
        self.done = [torch.cuda.Event() for _ in range(2)] if self.cuda else None
        self.free = [torch.cuda.Event() for _ in range(2)] if self.cuda else None
        self.free_recorded = [False, False]
        self.issued = self.consumed = 0

    def in_flight(self):
        return self.issued - self.consumed

    def issue(self, color=None, depth=None, packed=None):
        """Sources matter on the ingest rank only: either (color, depth) tensors or one already packed uint8 tensor."""
        import torch
        assert self.in_flight() < 2, "both slots hold frame sets that were not consumed"
        b = self.issued % 2
        if self.cuda:
            cur = torch.cuda.current_stream(self.device)
            if self.free_recorded[b]:
                cur.wait_event(self.free[b])
        if self.rank == self.src:
            if packed is not None:
                self.buf[b].copy_(packed, non_blocking=True)
            else:
                self.buf_color[b].copy_(color.reshape(-1).view(torch.uint8), non_blocking=True)
                self.buf_depth[b].copy_(depth.reshape(-1).view(torch.uint8), non_blocking=True)
        self.dist.broadcast(self.buf[b], self.src)
        if self.cuda:
            self.done[b].record(cur)
        self.issued += 1

    def consume(self, stream=None):
        assert self.in_flight() > 0, "no broadcast in flight"
        b = self.consumed % 2
        if self.cuda:
            stream.wait_event(self.done[b])
        self.consumed += 1
        return self.buf[b], b

    def release(self, slot, stream=None):
        if self.cuda:
            self.free[slot].record(stream)
            self.free_recorded[slot] = True

    def unpack(self, packed, color_shape, depth_shape):
        """Views of a packed frame set: (colour uint8 color_shape, depth float32 depth_shape)."""
        import torch
        return packed[:self.cb].view(color_shape), packed[self.cb:].view(torch.float32).view(depth_shape)


def gather_records(dist, records, dst=0, out=None):
    """The one gather per view: every rank's [h*w][8] float32 record image onto `dst` as [world][h*w][8]."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return records.unsqueeze(0)
    import torch
    if rank == dst:
        if out is None:
            out = torch.empty((world,) + tuple(records.shape), dtype=records.dtype, device=records.device)
        dist.gather(records, list(out.unbind(0)), dst=dst)
        return out
    dist.gather(records, None, dst=dst)
    return None


def reduce_records(dist, fu, records, keys, dst=0, stream=None):
    """Compositing by two reductions (rr_partial_keys / rr_partial_keep_winners): an all-reduce MIN of the int64 keys names
    every pixel's winner, every rank zeroes the records it lost, and an integer SUM reduce brings the winners' records to
    `dst` bit for bit. `records` [h*w][8] float32 and `keys` [h*w] int64 are device tensors of this rank; both are
    overwritten. On `dst`, `records` then holds the composited record image (rr_composite with n_parts = 1)."""
    import torch
    rank = dist.get_rank()
    cur = torch.cuda.current_stream(records.device) if records.is_cuda else None
    fu.partial_keys(records.data_ptr(), rank, keys.data_ptr())
    if cur is not None:
        cur.wait_stream(stream)
    dist.all_reduce(keys, op=dist.ReduceOp.MIN)
    if cur is not None:
        stream.wait_stream(cur)
    fu.partial_keep_winners(records.data_ptr(), keys.data_ptr(), rank)
    if cur is not None:
        cur.wait_stream(stream)
    dist.reduce(records.view(torch.int32), dst=dst, op=dist.ReduceOp.SUM)
    if cur is not None:
        stream.wait_stream(cur)
    return records


def composite_reference(parts: np.ndarray):
    """Pure-numpy statement of k_composite for tests of the gather layout: parts [P][n][8] -> (records [n][8], winner)."""
    steps = parts[..., 5].view(np.uint32) if parts.dtype == np.float32 else parts[..., 5]
    win = np.argmin(steps, axis=0)          # argmin returns the first (lowest rank) minimum, like the kernel
    return np.take_along_axis(parts, win[None, :, None], 0)[0], win
