"""Where does the end-to-end step time go? Times the ingest copy alone, the serial upload path and the pipelined
(stage/swap) path with and without the per-frame host sync. Diagnostic only; numbers printed here are not bench values."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rgbd-recon_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from rrpy import capi  # noqa: E402

scenes, inv, voxel = bench.make_inputs()
fu = capi.Fusion(bench.N_SENSORS, bench.W, bench.H, bench.CW, bench.CH, device=0)
capi.load_scene(fu, scenes[0], inv)
fu.configure(limit=bench.LIMIT, voxel_size=voxel, brick_size=bench.BRICK, min_voxels=bench.MIN_VOX, use_bricks=True)
hc = [torch.from_numpy(s.color).pin_memory() for s in scenes]
hd = [torch.from_numpy(s.depth).pin_memory() for s in scenes]
cb, db = hc[0].numel(), hd[0].numel() * 4
K = 200


def wall(fn, n=K):
    fn(0); fu.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    fu.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def copy_only(i):
    fu.stage_frames_ptr(hc[i % 2].data_ptr(), cb, hd[i % 2].data_ptr(), db)
    fu.stage_sync()
    fu.swap_frames()


def serial(i):
    fu.upload_frames_ptr(hc[i % 2].data_ptr(), cb, hd[i % 2].data_ptr(), db, device=False)
    fu.frame(sync_bricks=True)


def compute_only_sync(i):
    fu.frame(sync_bricks=True)


def compute_only_nosync(i):
    fu.frame(sync_bricks=False)


def pipelined(sync):
    def f(i):
        fu.swap_frames()
        fu.stage_frames_ptr(hc[(i + 1) % 2].data_ptr(), cb, hd[(i + 1) % 2].data_ptr(), db)
        fu.frame(sync_bricks=sync)
    return f


print("copy only           ms/frame", round(wall(copy_only), 4), " GB/s", round((cb + db) / wall(copy_only) / 1e6, 1))
print("compute, host sync  ms/frame", round(wall(compute_only_sync), 4))
print("compute, no sync    ms/frame", round(wall(compute_only_nosync), 4))
print("serial upload+frame ms/frame", round(wall(serial), 4))
fu.stage_frames_ptr(hc[0].data_ptr(), cb, hd[0].data_ptr(), db)
print("pipelined, sync     ms/frame", round(wall(pipelined(True)), 4))
print("pipelined, no sync  ms/frame", round(wall(pipelined(False)), 4))
fu.swap_frames()
fu.close()
