"""rr_view_export / rr_composite_peers (the group's peer-memory compositing with one process per GPU, as bench.py --gpus N uses
it): two PROCESSES, each owning one z-slab, march a view; the display process composites its own view with the other
process' view images opened through CUDA IPC. The result must equal the single-context view bit for bit. Both processes
may share one GPU (the test box has one); with two GPUs the second process takes the second one."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(ROOT, "rgbd-recon_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from rrpy import capi, synth
rank, world, tmp = int(sys.argv[1]), 2, sys.argv[2]
dev = rank % torch.cuda.device_count()
sc = synth.make_scene(N=2, W=128, H=106, CW=160, CH=136, cv_res=(32, 32, 64))
inv = synth.analytic_inverse(sc, (40, 44, 40))
mv = synth.look_at((1.4, 1.5, 2.0), (0.0, 1.1, 0.0)); vw, vh = 200, 120
pr = synth.perspective(50.0, vw / vh, 0.1, 10.0)
fu = capi.Fusion(sc.N, sc.W, sc.H, sc.CW, sc.CH, device=dev)
capi.load_scene(fu, sc, inv)
fu.configure(limit=0.01, voxel_size=0.02, brick_size=0.1, min_voxels=10, use_bricks=True)
Z = int(fu.volume_res()[2])
cut = Z // 2 + 3
fu.set_slab(0 if rank == 0 else cut, cut if rank == 0 else Z)
fu.upload_frames(sc.color, sc.depth)
fu.fuse_frame()
fu.raymarch(mv, pr, vw, vh, shade_mode=1, download=False)
fu.synchronize()
def wait_for(path):
    t0 = time.time()
    while not os.path.exists(path):
        assert time.time() - t0 < 120, "peer did not arrive: " + path
        time.sleep(0.01)
if rank == 1:
    open(os.path.join(tmp, "handle.tmp"), "wb").write(fu.view_export(vw, vh))
    os.rename(os.path.join(tmp, "handle.tmp"), os.path.join(tmp, "handle.bin"))      # the march is complete: rank 0 may read
    wait_for(os.path.join(tmp, "done"))                                               # keep the memory alive until rank 0 has composited
else:
    wait_for(os.path.join(tmp, "handle.bin"))
    rgba, depth = fu.composite_peers([open(os.path.join(tmp, "handle.bin"), "rb").read()], vw, vh)
    filled = fu.fill_colors()
    np.savez(os.path.join(tmp, "got.npz"), rgba=rgba, depth=depth, filled=filled)
    open(os.path.join(tmp, "done"), "w").write("1")
fu.close()
"""


def test_two_processes_composite_through_ipc(tmp_path):
    from rrpy import capi, synth
    env = dict(os.environ)
    code = "ROOT = %r\n" % ROOT + WORKER
    procs = [subprocess.Popen([sys.executable, "-c", code, str(r), str(tmp_path)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    got = np.load(str(tmp_path / "got.npz"))
    sc = synth.make_scene(N=2, W=128, H=106, CW=160, CH=136, cv_res=(32, 32, 64))
    inv = synth.analytic_inverse(sc, (40, 44, 40))
    mv = synth.look_at((1.4, 1.5, 2.0), (0.0, 1.1, 0.0)); vw, vh = 200, 120
    pr = synth.perspective(50.0, vw / vh, 0.1, 10.0)
    fu = capi.Fusion(sc.N, sc.W, sc.H, sc.CW, sc.CH)
    capi.load_scene(fu, sc, inv)
    fu.configure(limit=0.01, voxel_size=0.02, brick_size=0.1, min_voxels=10, use_bricks=True)
    fu.upload_frames(sc.color, sc.depth)
    fu.fuse_frame()
    rgba, depth = fu.raymarch(mv, pr, vw, vh, shade_mode=1)
    filled = fu.fill_colors()
    with pytest.raises(capi.RRError):
        fu.composite_peers([fu.view_export(vw, vh)], vw, vh)        # a handle of the same process cannot be opened
    fu.close()
    assert (depth < 1.0).sum() > 500
    assert np.array_equal(got["rgba"].view(np.uint32), rgba.view(np.uint32))
    assert np.array_equal(got["depth"].view(np.uint32), depth.view(np.uint32))
    assert np.array_equal(got["filled"].view(np.uint32), filled.view(np.uint32))
