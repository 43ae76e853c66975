"""GPU parity of the compressed stream formats: DXT1 colour blocks and 8-bit sqrt-compressed depth go
through rr_set_frame_format / rr_upload_frames and must give the stages and the volume the oracle computes from the
oracle-decoded layers (decoder pinned to the reference's external/squish in test_oracle_cpu.py)."""
import dataclasses

import numpy as np
import pytest

from conftest import bits_equal, mismatch_report

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def packed_scene():
    from rrpy import synth
    sc = synth.make_scene(N=2, W=128, H=106, CW=160, CH=136, cv_res=(32, 32, 64))     # colour size: multiples of 4
    dxt = np.stack([synth.encode_dxt1(sc.color[i]) for i in range(sc.N)])
    d8 = np.stack([synth.encode_depth8(sc.depth[i], 0.5, 4.5) for i in range(sc.N)])
    return sc, dxt, d8


def test_dxt5_ingest_matches_oracle(packed_scene):
    """compress_rgb == 5 (NetKinectArray.cpp:125-128,153-156): 16-byte blocks whose colour half is decoded in four-colour mode
    whatever the endpoint order (the swapped copy below decodes differently as DXT1); the alpha half is skipped."""
    import oracle_py as O
    from rrpy import capi, synth
    sc, _, _ = packed_scene
    inv = synth.analytic_inverse(sc, (40, 44, 40))
    blocks = []
    for i in range(sc.N):
        b = synth.encode_dxt5(sc.color[i], seed=i).reshape(-1, 16).copy()
        b[::3, 8:12] = b[::3, [10, 11, 8, 9]]           # every third block: endpoints swapped (c0 <= c1), indices kept
        blocks.append(b.reshape(-1))
    dxt5 = np.stack(blocks)
    fu = capi.Fusion(sc.N, sc.W, sc.H, sc.CW, sc.CH)
    capi.load_scene(fu, sc, inv)
    fu.configure(limit=0.01, voxel_size=0.025, brick_size=0.1, min_voxels=10, use_bricks=True)
    fu.set_frame_format(dxt5_color=True)
    with pytest.raises(capi.RRError):
        fu.upload_frames(dxt5[:, : dxt5.shape[1] // 2], sc.depth)      # DXT1-sized colour while DXT5 is expected
    fu.upload_frames(dxt5, sc.depth)
    fu.frame(sync_bricks=True)
    got_lab, got_tsdf = fu.download_stage("lab"), fu.download_tsdf()
    fu.close()
    color = np.stack([O.decode_dxt5(dxt5[i], sc.CW, sc.CH) for i in range(sc.N)])
    assert not np.array_equal(color[0], O.decode_dxt1(dxt5[0].reshape(-1, 16)[:, 8:].reshape(-1), sc.CW, sc.CH))
    osc = dataclasses.replace(sc, color=color)
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, 0.025, 0.1)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    pre = O.preprocess(osc, grid, cams, True, True, True)
    want = O.integrate(inv, pre, grid, 0.01, True, O.occupied_bricks(pre["bricks"], 10))
    assert bits_equal(got_lab, pre["lab"]).all(), mismatch_report("lab", got_lab, pre["lab"])
    assert bits_equal(got_tsdf, want).all(), mismatch_report("tsdf", got_tsdf, want)


@pytest.mark.parametrize("dxt1,depth8", [(True, False), (False, True), (True, True)])
def test_compressed_ingest_matches_oracle(packed_scene, dxt1, depth8):
    import oracle_py as O
    from rrpy import capi, synth
    sc, dxt, d8 = packed_scene
    inv = synth.analytic_inverse(sc, (40, 44, 40))
    near_far = np.float32([[0.5, 4.5]] * sc.N)
    fu = capi.Fusion(sc.N, sc.W, sc.H, sc.CW, sc.CH)
    capi.load_scene(fu, sc, inv)
    fu.configure(limit=0.01, voxel_size=0.025, brick_size=0.1, min_voxels=10, use_bricks=True)
    fu.set_frame_format(dxt1_color=dxt1, depth8=depth8, near_far=near_far if depth8 else None)
    fu.upload_frames(dxt if dxt1 else sc.color, d8 if depth8 else sc.depth)
    # reference quirk kept: pre_morph.fs validates metres (0.5 < d < 4.5) even when the texels are normalised bytes, so
    # 8-bit streams are only meaningful with useProcessedDepths(false)
    flags = (True, not depth8, True)
    fu.frame(*flags, sync_bricks=True)
    got = {k: fu.download_stage(k) for k in ("depth", "lab", "depth_b", "sil", "quality")}
    got["tsdf"] = fu.download_tsdf()
    counters, occupied = fu.download_bricks()
    fu.close()

    color = np.stack([O.decode_dxt1(dxt[i], sc.CW, sc.CH) for i in range(sc.N)]) if dxt1 else sc.color
    depth = O.depth8_to_float(d8) if depth8 else sc.depth
    osc = dataclasses.replace(sc, color=color, depth=depth)
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, 0.025, 0.1)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    pre = O.preprocess(osc, grid, cams, *flags, compress=near_far if depth8 else None)
    occ = O.occupied_bricks(pre["bricks"], 10)
    want = O.integrate(inv, pre, grid, 0.01, True, occ)
    assert len(occ) > 20
    for k in ("depth", "lab", "depth_b", "sil", "quality"):
        assert bits_equal(got[k], pre[k]).all(), mismatch_report(k, got[k], pre[k])
    assert np.array_equal(counters, pre["bricks"]) and np.array_equal(occupied, occ)
    assert bits_equal(got["tsdf"], want).all(), mismatch_report("tsdf", got["tsdf"], want)


def test_frame_format_errors(packed_scene):
    from rrpy import capi
    sc, dxt, d8 = packed_scene
    fu = capi.Fusion(sc.N, sc.W, sc.H, sc.CW, sc.CH)
    with pytest.raises(capi.RRError):
        fu.set_frame_format(depth8=True, near_far=None)            # 8-bit depth without its range
    fu.set_frame_format(dxt1_color=True)
    with pytest.raises(capi.RRError):
        fu.upload_frames(sc.color, sc.depth)                       # RGB8-sized colour while DXT1 is expected
    fu.close()
    odd = capi.Fusion(1, 64, 48, 66, 50)
    with pytest.raises(capi.RRError):
        odd.set_frame_format(dxt1_color=True)                      # DXT1 needs multiples of 4
    odd.close()
