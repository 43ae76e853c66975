"""GPU parity of the point-drawing reconstructions: rr_draw_points (ReconPoints::draw) and rr_draw_calibs
(ReconCalibs::draw) against the oracle's serial, draw-ordered rasterisation (oracle/ro_points.cpp), bit for bit, for every
shade mode, a view from inside the volume, 8 sensors and half2 voxels."""
import numpy as np
import pytest

from conftest import bits_equal, mismatch_report

pytestmark = pytest.mark.gpu


def _setup(sc, inv, voxel=0.02, fmt=0):
    import oracle_py as O
    from rrpy import capi
    fu = capi.Fusion(sc.N, sc.W, sc.H, sc.CW, sc.CH)
    capi.load_scene(fu, sc, inv)
    fu.configure(limit=0.01, voxel_size=voxel, brick_size=0.1, min_voxels=10, use_bricks=True, store_weight=fmt)
    fu.upload_frames(sc.color, sc.depth)
    fu.frame()
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, voxel, 0.1)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    pre = O.preprocess(sc, grid, cams)
    return fu, pre, grid


@pytest.mark.parametrize("eye", [(1.4, 1.5, 2.0), (0.2, 1.2, 0.3)])
def test_draw_points_matches_oracle(small_scene, eye):
    import oracle_py as O
    from rrpy import synth
    sc = small_scene
    inv = synth.analytic_inverse(sc, (40, 44, 40))
    fu, pre, _ = _setup(sc, inv)
    vw, vh = 240, 136
    mv, pr = synth.look_at(eye, (0.0, 1.1, 0.0)), synth.perspective(50.0, vw / vh, 0.1, 10.0)
    covered = 0
    for mode in range(4):
        rgba, depth = fu.draw_points(mv, pr, vw, vh, shade_mode=mode)
        w_rgba, w_depth = O.draw_points(sc, pre, mv, pr, vw, vh, shade_mode=mode)
        assert bits_equal(depth, w_depth).all(), mismatch_report(f"depth mode {mode}", depth, w_depth)
        assert bits_equal(rgba, w_rgba).all(), mismatch_report(f"rgba mode {mode}", rgba, w_rgba)
        covered = int((w_depth < 1.0).sum())
        assert (rgba[..., 3] == (depth < 1.0)).all()
    fu.close()
    assert covered > 1000                                    # the splats do cover the subject


def test_draw_points_eight_sensors():
    import oracle_py as O
    from rrpy import synth
    sc = synth.make_scene(N=8, W=96, H=80, CW=128, CH=108, cv_res=(24, 24, 48))
    inv = synth.analytic_inverse(sc, (40, 44, 40))
    fu, pre, _ = _setup(sc, inv, voxel=0.025)
    vw, vh = 200, 120
    mv, pr = synth.look_at((1.6, 1.4, -1.8), (0.0, 1.1, 0.0)), synth.perspective(45.0, vw / vh, 0.1, 10.0)
    for mode in (1, 3):
        rgba, depth = fu.draw_points(mv, pr, vw, vh, shade_mode=mode)
        w_rgba, w_depth = O.draw_points(sc, pre, mv, pr, vw, vh, shade_mode=mode)
        assert bits_equal(depth, w_depth).all() and bits_equal(rgba, w_rgba).all()
    fu.close()
    assert (w_depth < 1.0).sum() > 500


@pytest.mark.parametrize("fmt", [0, 2])
def test_draw_calibs_matches_oracle(small_scene, fmt):
    import oracle_py as O
    from rrpy import synth
    sc = small_scene
    inv = synth.analytic_inverse(sc, (40, 44, 40))
    fu, pre, grid = _setup(sc, inv, fmt=fmt)
    tsdf = fu.download_tsdf()                               # half2 voxels: the TSDF half, widened
    vw, vh = 240, 136
    mv, pr = synth.look_at((1.4, 1.5, 2.0), (0.0, 1.1, 0.0)), synth.perspective(50.0, vw / vh, 0.1, 10.0)
    seen = 0
    for limit in (0.01, 0.004):
        rgba, depth = fu.draw_calibs(mv, pr, vw, vh, active_kinect=1, limit=limit)
        w_rgba, w_depth = O.draw_calibs(tsdf, (40, 44, 40), limit, sc.bbox_min, sc.bbox_max, mv, pr, vw, vh)
        assert bits_equal(depth, w_depth).all(), mismatch_report("depth", depth, w_depth)
        assert bits_equal(rgba, w_rgba).all(), mismatch_report("rgba", rgba, w_rgba)
        seen = int((w_depth < 1.0).sum())
    from rrpy import capi
    with pytest.raises(capi.RRError):
        fu.draw_calibs(mv, pr, vw, vh, active_kinect=sc.N, limit=0.01)
    fu.close()
    assert seen > 50                                         # samples near the surface are not at -limit
