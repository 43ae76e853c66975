"""GPU parity of the raymarcher (rr_raymarch) against the oracle's restatement of tsdf_raymarch.fs.

Bar (BASELINE.json north_star): raymarched depth within 1 mm. The hit mask must be identical; positions are
compared in world space; colours / window depth within 1e-5. In practice the images are bit-identical."""
import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu

VW, VH = 320, 180


def _setup(scene, voxel, inv_res, use_bricks=True, skip_space=True):
    import oracle_py as O
    from rrpy import capi, synth
    inv = synth.analytic_inverse(scene, inv_res)
    fu = capi.Fusion(scene.N, scene.W, scene.H, scene.CW, scene.CH)
    capi.load_scene(fu, scene, inv)
    fu.configure(limit=0.01, voxel_size=voxel, brick_size=0.1, min_voxels=10, use_bricks=use_bricks, skip_space=skip_space)
    fu.upload_frames(scene.color, scene.depth)
    fu.frame(sync_bricks=True)
    grid = O.brick_grid(scene.bbox_min, scene.bbox_max, voxel, 0.1)
    cams = [O.frustum(scene.cv_xyz[i])[1] for i in range(scene.N)]
    pre = O.preprocess(scene, grid, cams)
    occ = O.occupied_bricks(pre["bricks"], 10)
    tsdf = O.integrate(inv, pre, grid, 0.01, use_bricks, occ)
    assert bits_equal(fu.download_tsdf(), tsdf).all()
    return fu, dict(inv=inv, grid=grid, pre=pre, occ=occ, tsdf=tsdf)


def _compare(scene, fu, st, eye, shade_mode, skip_space, use_bricks=True):
    import oracle_py as O
    from rrpy import synth
    mv = synth.look_at(eye, (0.0, 1.1, 0.0))
    pr = synth.perspective(50.0, VW / VH, 0.1, 10.0)
    rgba, depth = fu.raymarch(mv, pr, VW, VH, shade_mode=shade_mode)
    pos = fu.download_hit_positions(VW, VH)
    ns = fu.download_num_samples(VW, VH)
    want = O.raymarch(st["tsdf"], 0.01, st["inv"], scene, st["pre"], st["grid"], st["occ"], mv, pr, VW, VH, shade_mode,
                      skip_space=(skip_space and use_bricks))
    hit_w = want["depth"] < 1.0
    hit_g = pos[..., 3] > 0.5
    assert hit_w.sum() > 500, "view should see the surface"
    assert np.array_equal(hit_g, hit_w), f"hit masks differ on {(hit_g != hit_w).sum()} pixels"
    dims = (scene.bbox_max - scene.bbox_min).astype(np.float64)
    dmm = np.linalg.norm((pos[..., :3].astype(np.float64) - want["pos"]) * dims, axis=-1)[hit_w] * 1000.0
    assert dmm.max() <= 1.0, f"surface point differs by {dmm.max():.4f} mm (bar: 1 mm)"
    assert bits_equal(ns, want["samples"]).all(), "sample counts differ"
    ok = np.isfinite(want["rgba"]) & np.isfinite(rgba)
    assert (np.isnan(rgba) == np.isnan(want["rgba"])).all()
    assert np.abs(rgba[ok] - want["rgba"][ok]).max() <= 1e-5
    assert np.abs(depth - want["depth"]).max() <= 1e-6
    exact = bits_equal(rgba, want["rgba"]).all() and bits_equal(depth, want["depth"]).all() and bits_equal(pos[..., :3][hit_w], want["pos"][hit_w]).all()
    return exact, float(dmm.max())


@pytest.mark.parametrize("shade_mode", [0, 1, 2, 3])
def test_raymarch_skip_space(small_scene, shade_mode):
    fu, st = _setup(small_scene, 0.02, (50, 55, 50))
    exact, dmm = _compare(small_scene, fu, st, (1.6, 1.5, 2.2), shade_mode, True)
    fu.close()
    assert exact, f"image not bit-identical to the oracle (max surface deviation {dmm} mm)"


def test_raymarch_full_cube_march(small_scene):
    fu, st = _setup(small_scene, 0.02, (50, 55, 50), skip_space=False)
    exact, dmm = _compare(small_scene, fu, st, (-2.0, 1.0, 1.2), 0, False)
    fu.close()
    assert exact


def test_raymarch_dense_volume_camera_inside_box(small_scene):
    fu, st = _setup(small_scene, 0.025, (40, 44, 40), use_bricks=False)
    exact, dmm = _compare(small_scene, fu, st, (0.7, 1.3, 0.75), 1, True, use_bricks=False)
    fu.close()
    assert exact


def test_raymarch_half2_volume(small_scene):
    """Raymarch a half2-voxel volume (BASELINE config 5): must equal the oracle marching the fp16-rounded oracle volume."""
    import oracle_py as O
    from rrpy import capi, synth
    sc = small_scene
    inv = synth.analytic_inverse(sc, (50, 55, 50))
    fu = capi.Fusion(sc.N, sc.W, sc.H, sc.CW, sc.CH)
    capi.load_scene(fu, sc, inv)
    fu.configure(limit=0.01, voxel_size=0.02, brick_size=0.1, min_voxels=10, use_bricks=True, skip_space=True, store_weight=capi.VOXELS_HALF2)
    fu.upload_frames(sc.color, sc.depth)
    fu.frame(sync_bricks=True)
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, 0.02, 0.1)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    pre = O.preprocess(sc, grid, cams)
    occ = O.occupied_bricks(pre["bricks"], 10)
    tsdf = O.integrate(inv, pre, grid, 0.01, True, occ).astype(np.float16).astype(np.float32)
    assert bits_equal(fu.download_tsdf(), tsdf).all()
    st = dict(inv=inv, grid=grid, pre=pre, occ=occ, tsdf=tsdf)
    exact, dmm = _compare(sc, fu, st, (1.6, 1.5, 2.2), 1, True)
    fu.close()
    assert exact, f"image not bit-identical to the oracle on the rounded volume (max surface deviation {dmm} mm)"
