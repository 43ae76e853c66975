"""GPU parity of the triangle-mesh reconstruction: rr_draw_trigrid (ReconTrigrid::draw) against the oracle's
serial, draw-ordered rasterisation (oracle/ro_trigrid.cpp), bit for bit: every shade mode, a view from inside the volume
(triangles cut by the near plane), 8 sensors (several fragments per pixel: the additive blend's order matters), a fragment
pool that has to grow, and BASELINE's full size (4 x 512 x 424 sensors, 1280 x 720 view)."""
import numpy as np
import pytest

from conftest import bits_equal, mismatch_report

pytestmark = pytest.mark.gpu

MIN_LENGTH = 0.06        # the small scenes' depth pixels are ~2 cm apart: the reference's default 0.0125 (512 x 424) scaled up


def _setup(sc, inv, voxel=0.02):
    import oracle_py as O
    from rrpy import capi
    fu = capi.Fusion(sc.N, sc.W, sc.H, sc.CW, sc.CH)
    capi.load_scene(fu, sc, inv)
    fu.configure(limit=0.01, voxel_size=voxel, brick_size=0.1, min_voxels=10, use_bricks=True)
    fu.upload_frames(sc.color, sc.depth)
    fu.frame()
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, voxel, 0.1)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    return fu, O.preprocess(sc, grid, cams)


def _same(name, got, want):
    assert bits_equal(got[1], want[1]).all(), mismatch_report(f"{name} depth", got[1], want[1])
    assert bits_equal(got[0], want[0]).all(), mismatch_report(f"{name} rgba", got[0], want[0])


@pytest.mark.parametrize("eye", [(1.4, 1.5, 2.0), (0.2, 1.2, 0.3)])
def test_draw_trigrid_matches_oracle(small_scene, eye):
    import oracle_py as O
    from rrpy import synth
    sc = small_scene
    fu, pre = _setup(sc, synth.analytic_inverse(sc, (40, 44, 40)))
    vw, vh = 240, 136
    mv, pr = synth.look_at(eye, (0.0, 1.1, 0.0)), synth.perspective(50.0, vw / vh, 0.1, 10.0)
    covered = 0
    for mode in range(4):
        got = fu.draw_trigrid(mv, pr, vw, vh, shade_mode=mode, min_length=MIN_LENGTH)
        want = O.draw_trigrid(sc, pre, mv, pr, vw, vh, shade_mode=mode, min_length=MIN_LENGTH)
        _same(f"eye {eye} mode {mode}", got, want)
        covered = int((want[1] < 1.0).sum())
        assert (got[0][..., 3] == (got[1] < 1.0)).all()          # alpha 1 exactly where pass 1 left a surface that pass 2 kept
    # the same call again: the fragment lists are filled in whatever order the threads run, the sums must not care
    again = fu.draw_trigrid(mv, pr, vw, vh, shade_mode=3, min_length=MIN_LENGTH)
    _same("repeat", again, got)
    fu.close()
    assert covered > 400


def test_draw_trigrid_default_min_length_draws_nothing_here(small_scene):
    """validSurface rejects every triangle of the coarse test grid at the reference's default min_length: an all-background view."""
    from rrpy import synth
    sc = small_scene
    fu, _ = _setup(sc, synth.analytic_inverse(sc, (40, 44, 40)))
    mv, pr = synth.look_at((1.4, 1.5, 2.0), (0.0, 1.1, 0.0)), synth.perspective(50.0, 240 / 136, 0.1, 10.0)
    rgba, depth = fu.draw_trigrid(mv, pr, 240, 136)
    fu.close()
    assert (rgba == 0).all() and (depth == 1.0).all()


def test_draw_trigrid_eight_sensors_and_pool_growth():
    import oracle_py as O
    from rrpy import synth
    from rrpy import capi
    sc = synth.make_scene(N=8, W=96, H=80, CW=128, CH=108, cv_res=(24, 24, 48))
    capi.set_tunable("trigrid_pool", 1)          # room for one fragment per 16 pixels: the first view overflows it, pass 2 is repeated
    fu, pre = _setup(sc, synth.analytic_inverse(sc, (40, 44, 40)), voxel=0.025)
    blended = 0
    for (vw, vh), eye in (((200, 120), (1.6, 1.4, -1.8)), ((200, 120), (0.1, 1.3, 2.4)), ((320, 180), (-1.5, 1.2, 1.9))):
        mv, pr = synth.look_at(eye, (0.0, 1.1, 0.0)), synth.perspective(45.0, vw / vh, 0.1, 10.0)
        for mode in (1, 3):
            got = fu.draw_trigrid(mv, pr, vw, vh, shade_mode=mode, min_length=0.08)
            want = O.draw_trigrid(sc, pre, mv, pr, vw, vh, shade_mode=mode, min_length=0.08)
            _same(f"view {vw}x{vh} mode {mode}", got, want)
        # shade mode 3 paints every fragment with its sensor's colour: a pixel that is none of the eight pure colours was
        # blended from several sensors
        px = want[0][want[1] < 1.0][:, :3]
        pure = np.zeros(len(px), bool)
        for k in range(8):
            c = np.array([(228, 26, 28), (55, 126, 184), (77, 175, 74), (152, 78, 163), (255, 127, 0)][min(k, 4)], np.float32) / np.float32(255)
            pure |= (np.abs(px - c) < 1e-6).all(1)
        blended += int((~pure).sum())
    fu.close()
    capi.set_tunable("trigrid_pool", 64)
    assert blended > 50, "no pixel was blended from several sensors: the accumulation order is not exercised"


def test_draw_trigrid_fullsize():
    """BASELINE.json's sensors (4 x 512 x 424) into a 1280 x 720 view at the reference's default min_length."""
    import bench
    import oracle_py as O
    from rrpy import capi, synth
    scenes, inv, voxel = bench.make_inputs()
    sc = scenes[0]
    fu = capi.Fusion(bench.N_SENSORS, bench.W, bench.H, bench.CW, bench.CH)
    capi.load_scene(fu, sc, inv)
    fu.configure(limit=bench.LIMIT, voxel_size=voxel, brick_size=bench.BRICK, min_voxels=bench.MIN_VOX, use_bricks=True)
    fu.upload_frames(sc.color, sc.depth)
    fu.frame()
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, voxel, bench.BRICK)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    pre = O.preprocess(sc, grid, cams)
    vw, vh = 1280, 720
    mv, pr = synth.look_at((1.2, 1.5, 2.2), (0.0, 1.1, 0.0)), synth.perspective(50.0, vw / vh, 0.1, 10.0)
    got = fu.draw_trigrid(mv, pr, vw, vh, shade_mode=1)
    want = O.draw_trigrid(sc, pre, mv, pr, vw, vh, shade_mode=1)
    fu.close()
    _same("full size", got, want)
    assert (want[1] < 1.0).sum() > 20000
