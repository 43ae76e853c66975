"""BASELINE config 1: invert one synthetic 128x128x256 calibration volume (exact 8-NN + inverse-distance weighting +
frustum cull). Prints one JSON line: GPU seconds per volume / output Mvoxel/s, and the oracle port (the reference's own
parallelisation: OpenMP over x, calibration_inverter.cpp:121) timed on a bounded sample of output x-slices.
Not the headline bench (bench.py); numbers go to profiles/ and DESIGN.md."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "rgbd-recon_b200"), os.path.join(ROOT, "oracle")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--voxel", type=float, default=0.007)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from rrpy import capi, synth
    import oracle_py as O
    sc = synth.make_scene(N=1, W=512, H=424, CW=1280, CH=1080, cv_res=(128, 128, 256))
    res = tuple(int(np.ceil((sc.bbox_max[i] - sc.bbox_min[i]) / np.float32(a.voxel))) for i in range(3))
    fu = capi.Fusion(1, 512, 424, 1280, 1080)
    capi.load_scene(fu, sc)
    fu.set_timing(2)
    times = []
    for _ in range(a.reps + 1):
        fu.calib_invert(0, res, download=False)
        times.append(fu.stage_ms("calib_invert"))
    gpu_ms = float(np.median(times[1:]))
    got = fu.calib_invert(0, res)
    fu.close()
    nvox = res[0] * res[1] * res[2]
    # CPU: the same inversion restricted to a slab of the bounding box along x (the OpenMP-parallel axis), scaled
    cores = O.max_threads()
    nx = max(cores, int(res[0] * 0.04))
    x0 = res[0] // 2 - nx // 2
    sub_min, sub_max = sc.bbox_min.copy(), sc.bbox_max.copy()
    step = (sc.bbox_max[0] - sc.bbox_min[0]) / res[0]
    sub_min[0] = sc.bbox_min[0] + step * x0
    sub_max[0] = sc.bbox_min[0] + step * (x0 + nx)
    t0 = time.perf_counter()
    O.calib_invert(sc.cv_xyz[0], sub_min, sub_max, (nx, res[1], res[2]))
    cpu_s = time.perf_counter() - t0
    cpu_full = cpu_s * res[0] / nx
    valid = float((got[..., 3] > 0).mean())
    print(json.dumps({"workload": "calib_inverter: 128x128x256 cv_xyz -> inverse volume at ceil(bbox/%.3f)" % a.voxel, "out_res": res,
                      "valid_fraction": round(valid, 4), "gpu_ms_per_volume": round(gpu_ms, 3),
                      "gpu_mvoxel_per_s": round(nvox / gpu_ms / 1e3, 1),
                      "cpu": {"kind": "port", "cores": cores, "sample": f"{nx} of {res[0]} output x-slices ({cpu_s:.1f} s), scaled",
                              "seconds_per_volume": round(cpu_full, 2), "mvoxel_per_s": round(nvox / cpu_full / 1e6, 3)},
                      "speedup": round(cpu_full * 1e3 / gpu_ms, 1)}))


if __name__ == "__main__":
    main()
