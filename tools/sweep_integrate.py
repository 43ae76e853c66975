#!/usr/bin/env python
"""Sweep the integrator's launch-shape tunables on the bench workload (4 sensors, 512^3) on one GPU.

  python tools/sweep_integrate.py "threads=512,depth=2,chunk=1,zchunk=13" "threads=384,depth=4,chunk=0,zchunk=7" ...

For each configuration: 40 fused frames, mean CUDA-event time of the `2integrate` stage, and a check that the TSDF is
bit-identical to the first configuration's (tunables must never change results)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rgbd-recon_b200"))

import bench  # noqa: E402
from rrpy import capi  # noqa: E402

DEFAULTS = dict(fused=1, zchunk=13, fill_rows=16, fill_warps=2, ctas=2, threads=512, chunk=1, ldg256=1,
                staged=1, stage_zchunk=13, stage_ychunk=0, stage_tile=0, stage_fill_rows=16, stage_debug=0, stage_cwarps=0, stage_bulk_fill=4, stage_fill_depth=0, stage_fill_lsu=0, stage_tail_cap=2, fuse_nq=1)


def main():
    configs = sys.argv[1:] or ["threads=512"]
    scenes, inv, voxel = bench.make_inputs()
    fu = capi.Fusion(bench.N_SENSORS, bench.W, bench.H, bench.CW, bench.CH, device=0)
    capi.load_scene(fu, scenes[0], inv)
    fu.configure(limit=bench.LIMIT, voxel_size=voxel, brick_size=bench.BRICK, min_voxels=bench.MIN_VOX, use_bricks=True)
    fu.upload_frames(scenes[0].color, scenes[0].depth)
    ref_hash = None
    for spec in configs:
        kv = dict(DEFAULTS)
        for item in spec.split(","):
            if item:
                k, v = item.split("=")
                kv[k] = int(v)
        # non-tunable keys: min_voxels (occupancy threshold; huge = empty occupied list = pure clear), z0/z1 (slab)
        mv = kv.pop("min_voxels", bench.MIN_VOX)
        z0, z1 = kv.pop("z0", 0), kv.pop("z1", bench.R)
        for k, v in kv.items():
            capi.set_tunable(k, v)
        fu.configure(limit=bench.LIMIT, voxel_size=voxel, brick_size=bench.BRICK, min_voxels=mv, use_bricks=True)
        fu.set_slab(z0, z1)
        for _ in range(5):
            fu.frame()
        fu.synchronize()
        fu.set_timing(int(os.environ.get("SWEEP_TIMING", "1")))      # 2: per-pass timers too ("bricks")
        fu.stage_stats("2integrate"); fu.stage_stats("1preprocess"); fu.stage_stats("bricks")
        fu.integrator_profile()          # reset (counts only with the profiling build: RR_B200_LIB=.../librr_b200_prof.so, stage_debug bit 7)
        for _ in range(40):
            fu.frame()
        fu.synchronize()
        ms, n = fu.stage_stats("2integrate")
        pms, pn = fu.stage_stats("1preprocess")
        bms, bn = fu.stage_stats("bricks")
        fu.set_timing(0)
        h = hashlib.sha1(fu.download_tsdf().tobytes()).hexdigest()[:12]
        if ref_hash is None:
            ref_hash = h
        info = fu.integrator_info()
        info = {k: info[k] for k in ("staged", "tile", "zchunk", "oversize_pairs", "smem_bytes", "fill_warps", "flags", "slots", "slot_bytes")}
        rec = {"config": spec, "integrate_ms": round(ms / n, 5), "integrator": info, "preprocess_ms": round(pms / max(1, pn), 5), "bricks_ms": round(bms / max(1, bn), 5),
               "tsdf_sha1": h, "same_as_first": h == ref_hash}
        prof = fu.integrator_profile()
        if any(prof.values()):
            # per frame; cycle sums in kilocycles
            rec["roles_per_frame"] = {k: (v if k.endswith("max") else round(v / 40.0, 1) if k.endswith("items") else round(v / 40.0 / 1e3, 1)) for k, v in prof.items()}
        print(json.dumps(rec), flush=True)
    fu.close()


if __name__ == "__main__":
    main()
