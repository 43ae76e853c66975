#!/usr/bin/env python
"""Times one 1280x720 view of every renderer of the pre-processed maps on BASELINE's sensors (4 x 512 x 424, 512^3 volume):
rr_raymarch (+ rr_fill_colors), rr_draw_points, rr_draw_calibs, rr_draw_trigrid. Device time from the library's "draw" stage
timer (CUDA events around the kernels of one view), median of --views views. Prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "rgbd-recon_b200"), ROOT]

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=20)
    a = ap.parse_args()
    import bench
    from rrpy import capi, synth
    scenes, inv, voxel = bench.make_inputs()
    sc = scenes[0]
    fu = capi.Fusion(bench.N_SENSORS, bench.W, bench.H, bench.CW, bench.CH)
    capi.load_scene(fu, sc, inv)
    fu.configure(limit=bench.LIMIT, voxel_size=voxel, brick_size=bench.BRICK, min_voxels=bench.MIN_VOX, use_bricks=True)
    fu.upload_frames(sc.color, sc.depth)
    fu.frame(sync_bricks=True)
    fu.set_timing(2)
    vw, vh = 1280, 720
    mv, pr = synth.look_at((1.2, 1.5, 2.2), (0.0, 1.1, 0.0)), synth.perspective(50.0, vw / vh, 0.1, 10.0)
    out = {"view": [vw, vh], "sensors": [bench.N_SENSORS, bench.W, bench.H], "views": a.views}
    calls = {"raymarch": lambda: fu.raymarch(mv, pr, vw, vh, shade_mode=1),
             "draw_points": lambda: fu.draw_points(mv, pr, vw, vh, shade_mode=1),
             "draw_calibs": lambda: fu.draw_calibs(mv, pr, vw, vh, active_kinect=0, limit=bench.LIMIT),
             "draw_trigrid": lambda: fu.draw_trigrid(mv, pr, vw, vh, shade_mode=1)}
    for name, call in calls.items():
        ms = []
        l0 = fu.launch_count()
        fill = []
        for i in range(a.views + 2):
            r = call()
            if i >= 2:
                ms.append(fu.stage_ms("draw"))
            if name == "raymarch":                       # ReconIntegration::drawF: the colour hole filling follows the march
                fu.fill_colors(download=False)
                fu.synchronize()
                if i >= 2:
                    fill.append(fu.stage_ms("holefill"))
        if fill:
            out["fill_colors"] = {"ms_per_view": round(float(np.median(fill)), 4)}
        out[name] = {"ms_per_view": round(float(np.median(ms)), 4), "launches_per_view": (fu.launch_count() - l0) // (a.views + 2),
                     "covered_px": int((r[1] < 1.0).sum())}
    fu.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
