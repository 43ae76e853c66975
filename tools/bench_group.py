#!/usr/bin/env python
"""Single-process multi-GPU line of the bench workload through the product's own group API (rr_group_*: z-slabs, frame sets
by peer copies from the ingest device, per-slab integration, peer-memory compositing) - the C-ABI counterpart of
`bench.py --gpus N`, which drives one process per GPU over NCCL.

  python tools/bench_group.py --gpus N [--steps K] [--warmup W] [--c5]

Per step: rr_group_stage_frames of the NEXT frame set from pinned host memory (20 MB H2D + N-1 peer copies, overlapping the
kernels), rr_group_swap_frames, rr_group_fuse_frame. Timed with CUDA events on every member's stream, max over members.
Prints one JSON line (frames/s end to end, view ms, slab boundaries) and checks the fused volume and the composited view
against a single context bit for bit (`verified`)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rgbd-recon_b200"))

import torch  # noqa: E402

import bench  # noqa: E402
from rrpy import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--c5", action="store_true", help="BASELINE config 5: 8 sensors, 1024^3, half2 voxels")
    ap.add_argument("--devices", default="", help="explicit device list a,b,... (a device may repeat)")
    ap.add_argument("--dxt1", action="store_true", help="stream colour as DXT1 blocks and depth as 8-bit bytes (3.6 MB instead of 20 MB per frame set)")
    ap.add_argument("--count", action="store_true", help="read the occupied-brick count every step (the reference's per-frame read-back)")
    a = ap.parse_args()
    devices = [int(v) for v in a.devices.split(",")] if a.devices else list(range(a.gpus))
    n_s, res, fmt = (8, 1024, capi.VOXELS_HALF2) if a.c5 else (bench.N_SENSORS, bench.R, capi.VOXELS_F32)
    scenes, inv, voxel = bench.make_inputs(res, n_s)
    sc = scenes[0]

    def setup(obj):
        if isinstance(obj, capi.Group):
            obj.set_bbox(sc.bbox_min, sc.bbox_max)
            for i in range(sc.N):
                obj.calib_upload(i, sc.cv_xyz[i], sc.cv_uv[i])
                obj.calib_upload_inv(i, inv[i])
        else:
            capi.load_scene(obj, sc, inv)
        obj.configure(limit=bench.LIMIT, voxel_size=voxel, brick_size=bench.BRICK, min_voxels=bench.MIN_VOX, use_bricks=True, store_weight=fmt)

    near_far = np.float32([[0.5, 4.5]] * n_s)
    flags = (True, not a.dxt1, True)          # 8-bit depth streams skip pre_morph (it validates metres), as in the reference's use
    g = capi.Group(devices, n_s, bench.W, bench.H, bench.CW, bench.CH)
    setup(g)
    if a.dxt1:
        g.set_frame_format(dxt1_color=True, depth8=True, near_far=near_far)
        packed = [(np.stack([synth.encode_dxt1(s.color[i]) for i in range(n_s)]), np.stack([synth.encode_depth8(s.depth[i], 0.5, 4.5) for i in range(n_s)])) for s in scenes]
        pinned = [(torch.from_numpy(c).pin_memory(), torch.from_numpy(d).pin_memory()) for c, d in packed]
    else:
        pinned = [(torch.from_numpy(s.color).pin_memory(), torch.from_numpy(s.depth).pin_memory()) for s in scenes]
    streams = [torch.cuda.ExternalStream(g.member(i).stream(), device=torch.device("cuda", devices[i])) for i in range(len(devices))]

    def stage(k):
        c, d = pinned[k % len(pinned)]
        g.stage_frames_ptr(c.data_ptr(), c.numel() * c.element_size(), d.data_ptr(), d.numel() * d.element_size())

    def run(steps, first):
        for k in range(first, first + steps):
            g.swap_frames()
            stage(k + 1)
            g.fuse_frame(*flags)
            if a.count:
                g.bricks_count()
        return first + steps

    stage(0)
    k = run(2, 0)
    g.synchronize()
    g.balance_slabs()
    k = run(a.warmup, k)
    g.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in devices]
    for (e0, _), s in zip(ev, streams):
        e0.record(s)
    k = run(a.steps, k)
    for (_, e1), s in zip(ev, streams):
        e1.record(s)
    g.synchronize()
    ms = max(e0.elapsed_time(e1) for e0, e1 in ev) / a.steps
    n_occ, ratio = g.bricks_count()

    mv = synth.look_at((1.7, 1.6, 2.3), (0.0, 1.1, 0.0))
    pr = synth.perspective(50.0, bench.VW / bench.VH, 0.1, 10.0)
    for _ in range(3):
        g.raymarch(mv, pr, bench.VW, bench.VH, shade_mode=1, download=False); g.fill_colors(download=False)
    g.synchronize()
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    v0.record(streams[0])
    for _ in range(20):
        g.raymarch(mv, pr, bench.VW, bench.VH, shade_mode=1, download=False); g.fill_colors(download=False)
    v1.record(streams[0])
    g.synchronize()
    view_ms = v0.elapsed_time(v1) / 20

    # verification against one context on the ingest device: the frame set fused last, its volume and its view
    last = scenes[(k - 1) % len(scenes)]
    rgba, depth = g.raymarch(mv, pr, bench.VW, bench.VH, shade_mode=1)
    tsdf = g.download_tsdf()
    slabs = g.slabs()
    g.close()
    fu = capi.Fusion(n_s, bench.W, bench.H, bench.CW, bench.CH, device=devices[0])
    setup(fu)
    if a.dxt1:
        fu.set_frame_format(dxt1_color=True, depth8=True, near_far=near_far)
        c, d = packed[(k - 1) % len(scenes)]
        fu.upload_frames(c, d)
    else:
        fu.upload_frames(last.color, last.depth)
    fu.fuse_frame(*flags)
    w_tsdf = fu.download_tsdf()
    w_rgba, w_depth = fu.raymarch(mv, pr, bench.VW, bench.VH, shade_mode=1)
    fu.close()
    verified = bool(np.array_equal(tsdf.view(np.uint32), w_tsdf.view(np.uint32)) and np.array_equal(rgba.view(np.uint32), w_rgba.view(np.uint32)) and
                    np.array_equal(depth.view(np.uint32), w_depth.view(np.uint32)))
    print(json.dumps({"what": "rr_group (one process): stage (H2D + peer copies) + swap + fuse per frame set, end to end from pinned host memory",
                      "config": "c5: 8 sensors, 1024^3 half2" if a.c5 else bench.WORKLOAD, "devices": devices, "steps": a.steps, "warmup": a.warmup,
                      "ms_per_step": round(ms, 5), "frames_per_s": round(1e3 / ms, 2), "gvoxel_updates_per_s": round(res ** 3 / ms / 1e6, 3),
                      "h2d_bytes_per_step": int(pinned[0][0].numel() * pinned[0][0].element_size() + pinned[0][1].numel() * pinned[0][1].element_size()),
                      "stream": "DXT1 colour + 8-bit depth" if a.dxt1 else "RGB8 + float32", "view_ms": round(view_ms, 4),
                      "view": f"{bench.VW}x{bench.VH}: per-slab raymarch + peer-memory composite + colour hole filling",
                      "slabs": slabs, "occupied_bricks": n_occ, "verified": verified}), flush=True)


if __name__ == "__main__":
    main()
