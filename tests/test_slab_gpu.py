"""GPU parity of the z-slab path on ONE GPU: G slab contexts (as G ranks would hold them) integrate their slab + halo,
march their own samples into partial records, and the composite of the gathered records must equal the single-context
image bit for bit; each slab's owned TSDF slices must equal the full volume's."""
import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu
VW, VH = 320, 180


@pytest.mark.parametrize("G,use_bricks,shade", [(2, True, 1), (3, True, 0), (4, False, 3)])
def test_slab_composite_equals_single_volume(small_scene, G, use_bricks, shade):
    import torch
    from rrpy import capi, synth, multigpu as M
    sc = small_scene
    inv = synth.analytic_inverse(sc, (50, 55, 50))
    mv, pr = synth.look_at((1.6, 1.5, 2.2), (0.0, 1.1, 0.0)), synth.perspective(50.0, VW / VH, 0.1, 10.0)

    def make():
        fu = capi.Fusion(sc.N, sc.W, sc.H, sc.CW, sc.CH)
        capi.load_scene(fu, sc, inv)
        fu.configure(limit=0.01, voxel_size=0.02, brick_size=0.1, min_voxels=10, use_bricks=use_bricks)
        fu.upload_frames(sc.color, sc.depth)
        return fu

    full = make()
    full.frame(sync_bricks=True)
    want_rgba, want_depth = full.raymarch(mv, pr, VW, VH, shade_mode=shade)
    want_ns = full.download_num_samples(VW, VH)
    want_tsdf = full.download_tsdf()
    Z = want_tsdf.shape[0]
    assert (want_depth < 1).sum() > 500

    records = torch.empty((G, VW * VH, M.RECORD_FLOATS), dtype=torch.float32, device="cuda")
    slabs = []
    for r in range(G):
        fu = make()
        z0, z1 = M.slab_range(r, G, Z)
        fu.set_slab(z0, z1)
        fu.frame(sync_bricks=True)
        t = fu.download_tsdf()
        assert bits_equal(t[z0:z1], want_tsdf[z0:z1]).all(), f"slab {r} owned slices differ"
        h = M.halo(0.01, Z)
        lo, hi = max(0, z0 - h), min(Z, z1 + h)
        assert bits_equal(t[lo:hi], want_tsdf[lo:hi]).all(), f"slab {r} halo differs"
        fu.raymarch_partial(mv, pr, VW, VH, records[r].data_ptr(), shade_mode=shade)
        fu.synchronize()
        slabs.append(fu)
    rgba, depth = slabs[0].composite(records.data_ptr(), G, VW, VH)
    ns = slabs[0].download_num_samples(VW, VH)
    for fu in slabs:
        fu.close()
    full.close()
    assert bits_equal(depth, want_depth).all(), f"{(~bits_equal(depth, want_depth)).sum()} depth pixels differ"
    assert bits_equal(rgba, want_rgba).all(), f"{(~bits_equal(rgba, want_rgba)).sum()} colour values differ"
    hit = want_depth < 1
    assert bits_equal(ns[hit], want_ns[hit]).all(), "sample counts at hits differ"
