"""GPU parity of the colour hole filling (ReconIntegration::fillColors, recon_integration.cpp:280-339): the incremental
atlas kernels of csrc/rr_colorfill.cu against the oracle's literal simulation of the 2L-1 framebuffer passes."""
import numpy as np
import pytest

from conftest import bits_equal, mismatch_report

pytestmark = pytest.mark.gpu


def synthetic_view(W, H, seed, hole_frac=0.3):
    """A raymarch-like result: a blob of surface pixels, a fraction of them coloured by the fallback blend (alpha -1),
    some far-plane misses (depth >= 1 with alpha != 0) and background (alpha 0, depth 1)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    hit = ((xx - 0.45 * W) ** 2 / (0.38 * W) ** 2 + (yy - 0.5 * H) ** 2 / (0.42 * H) ** 2) < 1.0
    hit &= rng.random((H, W)) > 0.02
    rgba = np.zeros((H, W, 4), np.float32)
    depth = np.ones((H, W), np.float32)
    rgba[hit, :3] = rng.random((int(hit.sum()), 3), dtype=np.float32)
    holes = hit & (rng.random((H, W)) < hole_frac)
    big = (np.abs(xx - 0.5 * W) < 0.08 * W) & (np.abs(yy - 0.55 * H) < 0.1 * H)      # one large hole -> deep lods
    holes |= hit & big
    rgba[hit, 3] = 1.0
    rgba[holes, 3] = -1.0
    depth[hit] = (0.4 + 0.5 * rng.random(int(hit.sum()))).astype(np.float32)
    far = hit & (rng.random((H, W)) < 0.003)
    depth[far] = 1.0
    return rgba, depth


@pytest.mark.parametrize("W,H,seed", [(320, 180, 1), (321, 181, 2), (64, 48, 3), (1280, 720, 4), (100, 260, 5)])
def test_fill_colors_matches_oracle(W, H, seed):
    import oracle_py as O
    from rrpy import capi
    rgba, depth = synthetic_view(W, H, seed)
    want = O.fill_colors(rgba, depth)
    fu = capi.Fusion(1, 64, 48, 64, 48)
    fu.upload_view(rgba, depth)
    got = fu.fill_colors()
    launches = fu.launch_count()
    fu.close()
    assert bits_equal(got, want).all(), mismatch_report("filled colour", got, want)
    changed = ~bits_equal(got, rgba).all(-1)
    assert changed.sum() > 0 and launches <= 8
    # background pixels stay as they are, except where the shader's px / W * W texel arithmetic lands on a neighbour
    assert bits_equal(got[depth >= 1.0], rgba[depth >= 1.0]).all(-1).mean() > 0.98


def test_fill_after_raymarch(small_scene):
    import oracle_py as O
    from rrpy import capi, synth
    scene = small_scene
    inv = synth.analytic_inverse(scene, (40, 44, 40))
    fu = capi.Fusion(scene.N, scene.W, scene.H, scene.CW, scene.CH)
    capi.load_scene(fu, scene, inv)
    fu.configure(limit=0.01, voxel_size=0.025, brick_size=0.1, min_voxels=10, use_bricks=True)
    fu.upload_frames(scene.color, scene.depth)
    fu.frame(sync_bricks=True)
    mv = synth.look_at((1.2, 1.5, 1.9), (0.0, 1.1, 0.0))
    pr = synth.perspective(50.0, 320 / 200, 0.1, 10.0)
    rgba, depth = fu.raymarch(mv, pr, 320, 200, shade_mode=0)
    got = fu.fill_colors()
    fu.close()
    want = O.fill_colors(rgba, depth)
    assert (depth < 1.0).sum() > 1000
    assert bits_equal(got, want).all(), mismatch_report("filled colour after raymarch", got, want)
