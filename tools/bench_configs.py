#!/usr/bin/env python
"""Secondary throughput lines for the other BASELINE.json configs (bench.py measures the headline, config 4 at N = 1).

  python tools/bench_configs.py [c2] [c3] [c5]

  c2: 1 sensor, 128^3 TSDF integration + raymarch at 1280x720
  c3: 4 sensors, 256^3 TSDF with brick culling and colour fill
  c5: 8 sensors, 1024^3 TSDF, half2 (tsdf, weight) voxels (one GPU here; the slab path is bench.py --gpus N)

One JSON line per config: fused frames/s (device-resident inputs, CUDA events on the context's stream), stage times and
the view path. Parity of the same shapes at oracle-sized volumes is in tests/ (test_fusion_gpu.py, test_raymarch_gpu.py)."""
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

CONFIGS = {
    "c2": dict(N=1, R=128, fmt=capi.VOXELS_F32, what="1 sensor 512x424, 128^3 R32F TSDF integration + raymarch 1280x720"),
    "c3": dict(N=4, R=256, fmt=capi.VOXELS_F32, what="4 sensors, 256^3 R32F TSDF, brick culling, raymarch + colour fill"),
    "c5": dict(N=8, R=1024, fmt=capi.VOXELS_HALF2, what="8 sensors, 1024^3 TSDF, half2 (tsdf, weight) voxels, single GPU"),
}


def run(name, steps=100, warmup=5):
    cfg = CONFIGS[name]
    N, R = cfg["N"], cfg["R"]
    voxel = np.float32(bench.EXTENT / R)
    sc = synth.make_scene(N=N, W=bench.W, H=bench.H, CW=bench.CW, CH=bench.CH, cv_res=bench.CV_RES, bbox=bench.BBOX, frame=0, seed=1234)
    inv = synth.analytic_inverse(sc, bench.INV_RES)
    fu = capi.Fusion(N, bench.W, bench.H, bench.CW, bench.CH, device=0)
    capi.load_scene(fu, sc, inv)
    fu.configure(limit=bench.LIMIT, voxel_size=voxel, brick_size=bench.BRICK, min_voxels=bench.MIN_VOX, use_bricks=True, store_weight=cfg["fmt"])
    assert tuple(int(v) for v in fu.volume_res()) == (R, R, R)
    fu.upload_frames(sc.color, sc.depth)
    stream = torch.cuda.ExternalStream(fu.stream(), device=torch.device("cuda", 0))
    for _ in range(warmup):
        fu.frame(sync_bricks=False)
    fu.synchronize()
    fu.set_timing(1)
    fu.stage_stats("2integrate"); fu.stage_stats("1preprocess")
    fu.integrator_profile()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fu.frame(sync_bricks=False)
    e1.record(stream)
    fu.synchronize()
    ms = e0.elapsed_time(e1) / steps
    im, inn = fu.stage_stats("2integrate")
    pm, pn = fu.stage_stats("1preprocess")
    fu.set_timing(0)
    prof = fu.integrator_profile()       # counts only with the profiling build (RR_B200_LIB) and stage_debug bit 7
    n_occ, ratio = fu.bricks_update(sync=True)
    info = fu.integrator_info()
    VW, VH = 1280, 720
    mv = synth.look_at((1.7, 1.6, 2.3), (0.0, 1.1, 0.0))
    pr = synth.perspective(50.0, VW / VH, 0.1, 10.0)
    for _ in range(3):
        fu.raymarch(mv, pr, VW, VH, shade_mode=1, download=False); fu.fill_colors(download=False)
    fu.synchronize()
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    v0.record(stream)
    for _ in range(20):
        fu.raymarch(mv, pr, VW, VH, shade_mode=1, download=False); fu.fill_colors(download=False)
    v1.record(stream)
    fu.synchronize()
    fu.close()
    out = {"config": name, "what": cfg["what"], "frames_per_s": round(1e3 / ms, 2), "gvoxel_updates_per_s": round(R ** 3 / ms / 1e6, 3),
           "ms_per_step": round(ms, 5), "stages_ms": {"1preprocess": round(pm / max(1, pn), 5), "2integrate": round(im / max(1, inn), 5)},
           "occupied_bricks": int(n_occ), "occupied_ratio": round(float(ratio), 4), "integrator": info,
           "view_ms": round(v0.elapsed_time(v1) / 20, 4), "view": "raymarch 1280x720 (shaded, brick space skipping) + colour hole filling",
           "steps": steps, "warmup": warmup, "data": "synthetic (same generators as bench.py)"}
    if os.environ.get("RR_TUNE"):
        out["tunables"] = os.environ["RR_TUNE"]
    if any(prof.values()):
        out["roles_per_frame"] = {k: (v if k.endswith("max") else round(v / steps, 1) if k.endswith("items") else round(v / steps / 1e3, 1)) for k, v in prof.items()}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    # RR_TUNE="stage_debug=128,stage_tile=36": integrator tunables for A/B runs
    for kv in filter(None, os.environ.get("RR_TUNE", "").split(",")):
        k, v = kv.split("=")
        capi.set_tunable(k, int(v))
    for n in (sys.argv[1:] or ["c2", "c3", "c5"]):
        run(n)
