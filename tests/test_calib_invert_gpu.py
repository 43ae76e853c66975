"""GPU parity of rr_calib_invert (exact 8-NN + inverse-distance weighting + frustum cull) against the oracle, which is
itself pinned bit-for-bit to the reference's calibration_inverter.cpp (tests/test_oracle_cpu.py)."""
import os

import numpy as np
import pytest

from conftest import bits_equal, mismatch_report

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _invert(scene, sensor, res, keep=False):
    from rrpy import capi
    fu = capi.Fusion(scene.N, scene.W, scene.H, scene.CW, scene.CH)
    capi.load_scene(fu, scene)
    out = fu.calib_invert(sensor, res, keep=keep)
    return fu, out


def test_invert_matches_reference_golden():
    """Golden vector produced by the real CalibrationInverter (oracle/_ref)."""
    from rrpy import capi
    g = np.load(os.path.join(GOLD, "ref_calib_invert.npz"))
    xyz = g["cv_xyz"]
    Z, Y, X, _ = xyz.shape
    fu = capi.Fusion(1, 64, 53, 80, 68)
    fu.set_bbox(g["bbox_min"], g["bbox_max"])
    fu.calib_upload(0, xyz, np.zeros((Z, Y, X, 2), np.float32))
    got = fu.calib_invert(0, g["out_res"])
    assert bits_equal(fu.frustum_planes(0), g["planes"]).all()
    assert bits_equal(fu.camera_positions()[0], g["cam"]).all()
    fu.close()
    assert bits_equal(got, g["inv"]).all(), mismatch_report("cv_xyz_inv", got, g["inv"])


@pytest.mark.parametrize("res", [(40, 44, 40), (33, 21, 57)])
def test_invert_matches_oracle(small_scene, res):
    import oracle_py as O
    for sensor in range(small_scene.N):
        fu, got = _invert(small_scene, sensor, res)
        fu.close()
        want = O.calib_invert(small_scene.cv_xyz[sensor], small_scene.bbox_min, small_scene.bbox_max, res)
        assert 0.2 < (want[..., 3] > 0).mean() < 1.0
        assert bits_equal(got, want).all(), mismatch_report(f"cv_xyz_inv sensor {sensor}", got, want)


def test_invert_with_exact_distance_ties():
    """An undistorted lattice has many equidistant neighbours: ties are broken by sample index (x*Y*Z + y*Z + z)."""
    import oracle_py as O
    from rrpy import capi
    X, Y, Z = 9, 8, 10
    zz, yy, xx = np.meshgrid(np.arange(Z), np.arange(Y), np.arange(X), indexing="ij")
    xyz = np.stack([xx * 0.25 - 1.0, yy * 0.25, zz * 0.25 - 1.0], -1).astype(np.float32)
    bmin, bmax = np.array([-1, 0, -1], np.float32), np.array([1, 1.75, 1.25], np.float32)
    fu = capi.Fusion(1, 64, 53, 80, 68)
    fu.set_bbox(bmin, bmax)
    fu.calib_upload(0, xyz, np.zeros((Z, Y, X, 2), np.float32))
    got = fu.calib_invert(0, (16, 14, 18))       # voxel centres on lattice mid-points: 8-way ties
    fu.close()
    want = O.calib_invert(xyz, bmin, bmax, (16, 14, 18))
    assert bits_equal(got, want).all(), mismatch_report("cv_xyz_inv (ties)", got, want)


def test_inverted_volume_feeds_integration(small_scene):
    """rr_calib_invert(keep_on_device) output is usable as the sensor's inverse volume: integrate equals the oracle's."""
    import oracle_py as O
    from rrpy import capi
    res = (50, 55, 50)
    sc = small_scene
    fu = capi.Fusion(sc.N, sc.W, sc.H, sc.CW, sc.CH)
    capi.load_scene(fu, sc)
    inv = np.stack([fu.calib_invert(i, res, keep=True) for i in range(sc.N)])
    fu.configure(limit=0.01, voxel_size=0.02, brick_size=0.1, min_voxels=10, use_bricks=True)
    fu.upload_frames(sc.color, sc.depth)
    fu.frame(sync_bricks=True)
    tsdf = fu.download_tsdf()
    fu.close()
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, 0.02, 0.1)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    pre = O.preprocess(sc, grid, cams)
    occ = O.occupied_bricks(pre["bricks"], 10)
    want = O.integrate(inv, pre, grid, 0.01, True, occ)
    assert bits_equal(tsdf, want).all(), mismatch_report("tsdf", tsdf, want)
    assert ((want > -0.01) & (want < 0.01)).sum() > 1000


def test_inversion_at_baseline_size_matches_the_oracle():
    """BASELINE config 1 at its own size: one 128 x 128 x 256 calibration volume inverted over the default bounding box at
    7 mm (286 x 315 x 286 output voxels), every output voxel bit-identical to the oracle port (a few seconds on the host
    cores with the reference's own parallelisation)."""
    import oracle_py as O
    from rrpy import capi, synth
    sc = synth.make_scene(N=1, W=64, H=53, CW=80, CH=68, cv_res=(128, 128, 256))
    res = tuple(int(np.ceil((sc.bbox_max[i] - sc.bbox_min[i]) / np.float32(0.007))) for i in range(3))
    assert res == (286, 315, 286)
    fu = capi.Fusion(1, sc.W, sc.H, sc.CW, sc.CH)
    capi.load_scene(fu, sc)
    got = fu.calib_invert(0, res)
    fu.close()
    O.set_threads(max(1, len(os.sched_getaffinity(0))))
    want = O.calib_invert(sc.cv_xyz[0], sc.bbox_min, sc.bbox_max, res)
    assert (want[..., 3] > 0).mean() > 0.2
    assert bits_equal(got, want).all(), f"{(~bits_equal(got, want)).sum()} values differ"
