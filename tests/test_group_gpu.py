"""rr_group (one process, several GPUs, z-slabs; include/rgbd_recon_b200.h) against a single context: the fused volume
assembled from the members' slabs, the brick tables and the composited view must be bit-identical, for equal and for
cost-balanced slabs, with frames staged through the double-buffered peer-copy path. Runs on ONE GPU by putting several
members on the same device (the slab, peer-copy and peer-composite code is the same); uses distinct devices when the box
has them."""
import numpy as np
import pytest

from conftest import bits_equal, mismatch_report

pytestmark = pytest.mark.gpu


def _devices(n):
    import torch
    have = torch.cuda.device_count()
    return [i % max(1, have) for i in range(n)]


def _single(sc, inv, voxel, mv, pr, vw, vh, frames):
    from rrpy import capi
    fu = capi.Fusion(sc.N, sc.W, sc.H, sc.CW, sc.CH)
    capi.load_scene(fu, sc, inv)
    fu.configure(limit=0.01, voxel_size=voxel, brick_size=0.1, min_voxels=10, use_bricks=True)
    out = []
    for s in frames:
        fu.upload_frames(s.color, s.depth)
        fu.fuse_frame()
        n_occ, _ = fu.bricks_count()
        tsdf = fu.download_tsdf()
        rgba, depth = fu.raymarch(mv, pr, vw, vh, shade_mode=1)
        filled = fu.fill_colors()
        out.append((n_occ, tsdf, rgba, depth, filled))
    fu.close()
    return out


@pytest.mark.parametrize("n", [1, 2, 3, 5])
def test_group_equals_single_context(n):
    from rrpy import capi, synth
    sc = synth.make_scene(N=2, W=128, H=106, CW=160, CH=136, cv_res=(32, 32, 64))
    frames = [sc, synth.rerender(sc, 5)]
    inv = synth.analytic_inverse(sc, (40, 44, 40))
    voxel = 0.02
    mv = synth.look_at((1.4, 1.5, 2.0), (0.0, 1.1, 0.0))
    vw, vh = 200, 120
    pr = synth.perspective(50.0, vw / vh, 0.1, 10.0)
    want = _single(sc, inv, voxel, mv, pr, vw, vh, frames)

    g = capi.Group(_devices(n), sc.N, sc.W, sc.H, sc.CW, sc.CH)
    g.set_bbox(sc.bbox_min, sc.bbox_max)
    for i in range(sc.N):
        g.calib_upload(i, sc.cv_xyz[i], sc.cv_uv[i])
        g.calib_upload_inv(i, inv[i])
    g.configure(limit=0.01, voxel_size=voxel, brick_size=0.1, min_voxels=10, use_bricks=True)
    Z = int(g.member(0).volume_res()[2])
    assert g.slabs()[0] == 0 and g.slabs()[-1] == Z and len(g.slabs()) == n + 1
    for k, s in enumerate(frames):
        g.upload_frames(s.color, s.depth)
        if k == 0:
            g.fuse_frame()
        else:
            g.frame()                                  # the call-by-call form on every member
        n_occ, _ = g.bricks_count()
        tsdf = g.download_tsdf()
        rgba, depth = g.raymarch(mv, pr, vw, vh, shade_mode=1)
        filled = g.fill_colors()
        w_occ, w_tsdf, w_rgba, w_depth, w_filled = want[k]
        assert n_occ == w_occ and n_occ > 10
        assert bits_equal(tsdf, w_tsdf).all(), mismatch_report(f"tsdf frame {k}", tsdf, w_tsdf)
        assert bits_equal(rgba, w_rgba).all(), mismatch_report(f"rgba frame {k}", rgba, w_rgba)
        assert bits_equal(depth, w_depth).all(), mismatch_report(f"depth frame {k}", depth, w_depth)
        assert bits_equal(filled, w_filled).all(), mismatch_report(f"filled frame {k}", filled, w_filled)
        assert (depth < 1.0).sum() > 500                # the view does see the surface
        if k == 0 and n > 1:
            g.balance_slabs()                           # equal-cost slabs from this frame's occupied bricks
            b = g.slabs()
            assert b[0] == 0 and b[-1] == Z and all(b[i] < b[i + 1] for i in range(n))
    g.close()


def test_group_pipelined_staging_and_compressed_streams():
    """Frame sets staged one ahead (rr_group_stage_frames while the previous set is fused), in the DXT1 + 8-bit stream format:
    every fused volume equals the single-context volume of the same frame set."""
    import torch
    from rrpy import capi, synth
    sc = synth.make_scene(N=2, W=128, H=106, CW=160, CH=136, cv_res=(32, 32, 64))
    sets = [sc] + [synth.rerender(sc, 3 * t) for t in range(1, 4)]
    inv = synth.analytic_inverse(sc, (40, 44, 40))
    near_far = np.float32([[0.5, 4.5]] * sc.N)
    packed = []
    for s in sets:
        dxt = np.stack([synth.encode_dxt1(s.color[i]) for i in range(s.N)])
        d8 = np.stack([synth.encode_depth8(s.depth[i], 0.5, 4.5) for i in range(s.N)])
        packed.append((torch.from_numpy(dxt).pin_memory(), torch.from_numpy(d8).pin_memory()))

    def volumes(obj):
        obj.set_frame_format(dxt1_color=True, depth8=True, near_far=near_far)
        out = []
        c, d = packed[0]
        obj.stage_frames_ptr(c.data_ptr(), c.numel(), d.data_ptr(), d.numel())
        for k in range(len(packed)):
            obj.swap_frames()
            if k + 1 < len(packed):
                c, d = packed[k + 1]
                obj.stage_frames_ptr(c.data_ptr(), c.numel(), d.data_ptr(), d.numel())      # overlaps the fuse below
            obj.fuse_frame(True, False, True)
            out.append(obj.download_tsdf())
        return out

    fu = capi.Fusion(sc.N, sc.W, sc.H, sc.CW, sc.CH)
    capi.load_scene(fu, sc, inv)
    fu.configure(limit=0.01, voxel_size=0.02, brick_size=0.1, min_voxels=10, use_bricks=True)
    want = volumes(fu)
    fu.close()
    g = capi.Group(_devices(4), sc.N, sc.W, sc.H, sc.CW, sc.CH)          # four members: the frame sets go down a two-level tree
    g.set_bbox(sc.bbox_min, sc.bbox_max)
    for i in range(sc.N):
        g.calib_upload(i, sc.cv_xyz[i], sc.cv_uv[i])
        g.calib_upload_inv(i, inv[i])
    g.configure(limit=0.01, voxel_size=0.02, brick_size=0.1, min_voxels=10, use_bricks=True)
    got = volumes(g)
    g.close()
    assert len({w.tobytes() for w in want}) == len(want)          # the frame sets do differ
    for k in range(len(want)):
        assert bits_equal(got[k], want[k]).all(), mismatch_report(f"tsdf of set {k}", got[k], want[k])


def test_group_errors():
    from rrpy import capi
    with pytest.raises(capi.RRError):
        capi.Group([], 1, 64, 48, 64, 48)
    g = capi.Group(_devices(2), 1, 64, 48, 64, 48)
    with pytest.raises(capi.RRError, match="member 0"):
        g.configure()                                  # no bbox yet: the member's message comes through
    g.close()
