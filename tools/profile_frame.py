"""Runs a few fused frames of the bench workload (for ncu captures; numbers printed here are never bench values)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rgbd-recon_b200"))
import bench  # noqa: E402
from rrpy import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="bricks")
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--res", type=int, default=bench.R)
a = ap.parse_args()
scenes, inv, voxel = bench.make_inputs(a.res)
fu = capi.Fusion(bench.N_SENSORS, bench.W, bench.H, bench.CW, bench.CH)
capi.load_scene(fu, scenes[0], inv)
fu.configure(limit=bench.LIMIT, voxel_size=voxel, brick_size=bench.BRICK, min_voxels=bench.MIN_VOX, use_bricks=(a.mode == "bricks"))
for i in range(a.frames):
    s = scenes[i % len(scenes)]
    fu.upload_frames(s.color, s.depth)
    print(i, fu.frame(sync_bricks=True))
# one view of every renderer, so that launch lists and captures also hold the view-path kernels
from rrpy import synth  # noqa: E402
mv, pr = synth.look_at((1.7, 1.6, 2.3), (0.0, 1.1, 0.0)), synth.perspective(50.0, bench.VW / bench.VH, 0.1, 10.0)
fu.raymarch(mv, pr, bench.VW, bench.VH, shade_mode=1, download=False)
fu.fill_colors(download=False)
fu.draw_points(mv, pr, bench.VW, bench.VH, shade_mode=1)
fu.draw_calibs(mv, pr, bench.VW, bench.VH)
fu.synchronize()
fu.close()
