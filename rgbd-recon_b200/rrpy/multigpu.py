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


def halo(limit: float, Z: int) -> int:
    """Slices a slab owner recomputes past its slab on each side (rr_integrate does this internally)."""
    return int(np.ceil(np.float32(limit) * np.float32(Z))) + 2


def broadcast_frames(dist, color, depth, src=0):
    """One frame set (colour uint8 [N][CH][CW][3], depth float32 [N][H][W]) from the ingest rank to every rank.
    Tensors must be preallocated on every rank; returns them for convenience."""
    dist.broadcast(color, src)
    dist.broadcast(depth, src)
    return color, depth


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


def composite_reference(parts: np.ndarray):
    """Pure-numpy statement of k_composite for tests of the gather layout: parts [P][n][8] -> (records [n][8], winner)."""
    steps = parts[..., 5].view(np.uint32) if parts.dtype == np.float32 else parts[..., 5]
    win = np.argmin(steps, axis=0)          # argmin returns the first (lowest rank) minimum, like the kernel
    return np.take_along_axis(parts, win[None, :, None], 0)[0], win
