"""Generates tests/golden/*.npz. Run in the build container (needs /root/reference for oracle/_ref):

    make -C oracle all && python tools/make_golden.py

ref_*.npz hold outputs of REAL reference code (oracle/_ref, see oracle/ref_harness.cpp) on seeded inputs; they travel to
machines without /root/reference. oracle_*.npz are self-regression pins of the scalar restatement (the reference's GLSL
cannot run here), so that a later edit of oracle/*.cpp that changes results is caught."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "rgbd-recon_b200")]
import oracle_py as O  # noqa: E402
import ref_py as R  # noqa: E402
from rrpy import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_scene():
    return synth.make_scene(N=1, W=64, H=53, CW=80, CH=68, cv_res=(16, 16, 32), seed=77)


def glsl_scene():
    # large enough for the 13x13 support tests of pre_depth / pre_boundary to keep a surface (golden_scene() is too small)
    return synth.make_scene(N=1, W=128, H=106, CW=160, CH=135, cv_res=(24, 24, 48), seed=77)


def golden_dxt1():
    """DXT1 colour ingest: blocks made and decoded by the reference's own codec external/squish."""
    rng = np.random.default_rng(11)
    H, W = 48, 64
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.stack([(xx * 4) % 256, (yy * 5) % 256, ((xx // 8 + yy // 8) % 2) * 200 + 20], axis=2).astype(np.int32)
    img = np.clip(img + rng.integers(-12, 13, size=img.shape), 0, 255).astype(np.uint8)
    blocks = R.squish_compress_dxt1(img)
    decoded = R.squish_decompress_dxt1(blocks, W, H)
    # arbitrary block bytes exercise both endpoint orders (4-colour and 3-colour + transparent-black mode)
    rnd = rng.integers(0, 256, size=(W // 4) * (H // 4) * 8, dtype=np.uint8)
    rnd_decoded = R.squish_decompress_dxt1(rnd, W, H)
    np.savez_compressed(os.path.join(OUT, "ref_dxt1.npz"), image=img, blocks=blocks, decoded=decoded, random_blocks=rnd,
                        random_decoded=rnd_decoded)


def golden_dxt5():
    """DXT5 colour ingest (compress_rgb == 5): blocks made and decoded by external/squish, plus arbitrary block bytes."""
    rng = np.random.default_rng(12)
    H, W = 48, 64
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.stack([(xx * 3 + yy) % 256, (yy * 5) % 256, ((xx // 4 + yy // 8) % 2) * 180 + 40], axis=2).astype(np.int32)
    img = np.clip(img + rng.integers(-12, 13, size=img.shape), 0, 255).astype(np.uint8)
    alpha = rng.integers(0, 256, size=(H, W), dtype=np.uint8)
    blocks = R.squish_compress_dxt5(img, alpha)
    decoded = R.squish_decompress_dxt5(blocks, W, H)
    rnd = rng.integers(0, 256, size=(W // 4) * (H // 4) * 16, dtype=np.uint8)
    rnd_decoded = R.squish_decompress_dxt5(rnd, W, H)
    np.savez_compressed(os.path.join(OUT, "ref_dxt5.npz"), image=img, blocks=blocks, decoded=decoded, random_blocks=rnd,
                        random_decoded=rnd_decoded)


def golden_glsl():
    """The reference's OWN shaders (glsl/pre_*.fs, inc_*.glsl, tsdf_integration.vs) compiled as C++ and run on the CPU
    (oracle/_ref/libref_glsl.so, oracle/glsl_host/): every pre-processing stage and the integrated volume on the golden
    scene, chained exactly like NetKinectArray::processTextures + ReconIntegration::integrate."""
    import ref_glsl_py as G
    assert G.available(), "build oracle/_ref/libref_glsl.so first (make -C oracle all)"
    sc = glsl_scene()
    voxel = 0.035
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, voxel, 0.1)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    pre = G.preprocess(sc, grid, cams)
    inv = synth.analytic_inverse(sc, (20, 22, 20))
    occ = O.occupied_bricks(pre["bricks"], 10)
    tsdf_bricks = G.integrate(inv, pre, grid, 0.01, True, occ)
    tsdf_dense = G.integrate(inv, pre, grid, 0.01, False, occ)
    np.savez_compressed(os.path.join(OUT, "ref_glsl_stages.npz"), voxel=np.float32(voxel), inv=inv, occupied=occ,
                        tsdf_bricks=tsdf_bricks, tsdf_dense=tsdf_dense, **{"pre_" + k: v for k, v in pre.items()})


RM_VIEW = dict(eye=(1.2, 1.4, 1.6), at=(0.0, 1.1, 0.0), fovy=50.0, w=160, h=90)


def golden_glsl_raymarch():
    """glsl/tsdf_raymarch.fs + shading.glsl run on the CPU (cube-proxy march, skipSpace off) on the ORACLE's stages and volume
    of glsl_scene(): inputs every machine can regenerate bit for bit, outputs of the reference's shader."""
    import ref_glsl_py as G
    sc = glsl_scene()
    voxel = 0.035
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, voxel, 0.1)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    pre = O.preprocess(sc, grid, cams)
    inv = synth.analytic_inverse(sc, (20, 22, 20))
    occ = O.occupied_bricks(pre["bricks"], 10)
    tsdf = O.integrate(inv, pre, grid, 0.01, True, occ)
    mv = synth.look_at(RM_VIEW["eye"], RM_VIEW["at"])
    pr = synth.perspective(RM_VIEW["fovy"], RM_VIEW["w"] / RM_VIEW["h"], 0.1, 10.0)
    out = {}
    for mode in range(4):
        r = G.raymarch(tsdf, 0.01, inv, sc, pre, mv, pr, RM_VIEW["w"], RM_VIEW["h"], mode)
        out[f"rgba{mode}"] = r["rgba"]
        out["depth"], out["samples"], out["hit"] = r["depth"], r["samples"], r["hit"]
    # the skipSpace branch (getStartPos / screenToVol) on the depth peels of drawDepthLimits (the reference's brick shaders)
    peels = G.depth_peels_rasterised(sc, grid, pre["bricks"], occ, mv, pr, RM_VIEW["w"], RM_VIEW["h"])    # bricks.vs/gs/fs + rasteriser
    r = G.raymarch(tsdf, 0.01, inv, sc, pre, mv, pr, RM_VIEW["w"], RM_VIEW["h"], 1, depth_peels=peels)
    out.update(skip_rgba=r["rgba"], skip_depth=r["depth"], skip_samples=r["samples"], skip_hit=r["hit"])
    np.savez_compressed(os.path.join(OUT, "ref_glsl_raymarch.npz"), voxel=np.float32(voxel), tsdf_sha=np.array(sha(tsdf)), **out)
    # colour hole filling (framebuffer_transfer.fs, tsdf_inpaint.fs, tsdf_colorfill.fs) on the shaded image above
    filled, atlas_c, atlas_d = G.fill_colors(out["rgba1"], out["depth"], want_atlas=True)
    np.savez_compressed(os.path.join(OUT, "ref_glsl_colorfill.npz"), rgba=out["rgba1"], depth=out["depth"], filled=filled,
                        atlas_color=atlas_c, atlas_depth=atlas_d)


def golden_glsl_points():
    """glsl/points.{vs,gs,fs} and calib_vis.{vs,fs} run on the CPU (oracle/glsl_host: ReconPoints::draw, ReconCalibs::draw with the
    fixed-function point pipeline of OpenGL 4.4) on the ORACLE's stages and volume of glsl_scene()."""
    import ref_glsl_py as G
    sc = glsl_scene()
    voxel = 0.035
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, voxel, 0.1)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    pre = O.preprocess(sc, grid, cams)
    inv = synth.analytic_inverse(sc, (20, 22, 20))
    tsdf = O.integrate(inv, pre, grid, 0.01, True, O.occupied_bricks(pre["bricks"], 10))
    mv = synth.look_at(RM_VIEW["eye"], RM_VIEW["at"])
    pr = synth.perspective(RM_VIEW["fovy"], RM_VIEW["w"] / RM_VIEW["h"], 0.1, 10.0)
    out = {}
    for mode in range(4):
        out[f"points_rgba{mode}"], out[f"points_depth{mode}"] = G.draw_points(sc, pre, mv, pr, RM_VIEW["w"], RM_VIEW["h"], mode)
    out["calibs_rgba"], out["calibs_depth"] = G.draw_calibs(tsdf, inv, sc, 0, 0.01, mv, pr, RM_VIEW["w"], RM_VIEW["h"])
    np.savez_compressed(os.path.join(OUT, "ref_glsl_points.npz"), voxel=np.float32(voxel), tsdf_sha=np.array(sha(tsdf)), **out)


TRIGRID_MIN_LENGTH = 0.06     # the golden scene's depth pixels are ~2 cm apart (the reference's 0.0125 belongs to 512 x 424 sensors)


def golden_glsl_trigrid():
    """glsl/trigrid_accum.{vs,gs,fs} and trigrid_normalize.fs run on the CPU (oracle/glsl_host: ReconTrigrid::draw through the
    OpenGL 4.4 fixed-function stages of oracle/ro_raster.h) on the ORACLE's stages of glsl_scene(), every shade mode, the
    camera outside (RM_VIEW) and inside the volume (triangles cut by the near plane)."""
    import ref_glsl_py as G
    sc = glsl_scene()
    voxel = 0.035
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, voxel, 0.1)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    pre = O.preprocess(sc, grid, cams)
    pr = synth.perspective(RM_VIEW["fovy"], RM_VIEW["w"] / RM_VIEW["h"], 0.1, 10.0)
    out = {}
    for tag, eye in (("", RM_VIEW["eye"]), ("in_", (0.3, 1.2, 0.4))):
        mv = synth.look_at(eye, RM_VIEW["at"])
        for mode in range(4):
            out[f"{tag}rgba{mode}"], out[f"{tag}depth{mode}"] = G.draw_trigrid(sc, pre, mv, pr, RM_VIEW["w"], RM_VIEW["h"], mode, TRIGRID_MIN_LENGTH)
    np.savez_compressed(os.path.join(OUT, "ref_glsl_trigrid.npz"), voxel=np.float32(voxel), min_length=np.float32(TRIGRID_MIN_LENGTH),
                        quality_sha=np.array(sha(pre["quality"])), eye_in=np.array((0.3, 1.2, 0.4), np.float32), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    assert R.available(), "build oracle/_ref first (make -C oracle all)"
    if "--only-trigrid" in sys.argv:
        golden_glsl_trigrid()
        return
    if "--only-points" in sys.argv:
        golden_glsl_points()
        return
    if "--only-dxt" in sys.argv:
        golden_dxt1()
        golden_dxt5()
        return
    if "--only-glsl" in sys.argv:
        golden_glsl()
        golden_glsl_raymarch()
        return
    golden_dxt1()
    golden_dxt5()
    golden_glsl_points()
    golden_glsl_trigrid()
    golden_glsl()
    golden_glsl_raymarch()
    sc = golden_scene()
    xyz = sc.cv_xyz[0]
    # --- calibration inversion + frustum (real calibration_inverter.cpp / frustum.cpp)
    res = (20, 22, 20)
    inv = R.calib_invert(xyz, sc.bbox_min, sc.bbox_max, res)
    planes, cam = R.frustum(xyz)
    rng = np.random.default_rng(5)
    pts = rng.uniform(-2.5, 2.5, size=(4096, 3)).astype(np.float32) + np.array([0, 1.1, 0], np.float32)
    inside = R.frustum_inside(xyz, pts)
    np.savez_compressed(os.path.join(OUT, "ref_calib_invert.npz"), cv_xyz=xyz, bbox_min=sc.bbox_min, bbox_max=sc.bbox_max,
                        out_res=np.array(res, np.uint32), inv=inv, planes=planes, cam=cam, points=pts, inside=inside)
    # --- volume file format (real calibration_volume.hpp)
    small = rng.normal(size=(3, 4, 5, 4)).astype(np.float32)
    back, res_out, lim_out, raw = R.volume_roundtrip(small, (0.5, 4.5))
    np.savez_compressed(os.path.join(OUT, "ref_volume_file.npz"), data=small, raw=np.frombuffer(raw, np.uint8), back=back)
    # --- brick membership (real volume_sampler.cpp containedVoxels) + voxel centres + glm::round
    cases = []
    for dims, voxel, brick in (((20, 22, 20), 0.1, 0.3), ((50, 55, 50), 0.04, 0.1), ((67, 74, 67), 0.03, 0.1), ((29, 31, 23), 0.07, 0.21)):
        g = O.brick_grid(np.zeros(3, np.float32), (np.array(dims) * voxel).astype(np.float32), voxel, brick)
        # the same pos/size arguments divideBox passes (recon_integration.cpp:367-388), for a sample of bricks
        cases.append((tuple(int(v) for v in g["res"]), voxel, brick, g, (np.array(dims) * voxel).astype(np.float32)))
    cv = {}
    for ci, (dims, voxel, brick, g, bmax) in enumerate(cases):
        args = O.divide_box_args(np.zeros(3, np.float32), bmax, g["brick_size"], res=g["res"])
        sel = np.unique(np.concatenate([np.arange(0, len(args), max(1, len(args) // 40)), [len(args) - 1]]))
        lists = []
        for b in sel:
            idx, n = R.contained_voxels(dims, args[b, :3], args[b, 3:])
            assert n == len(idx)
            lists.append(idx)
        cv[f"c{ci}_dims"] = np.array(dims, np.uint32)
        cv[f"c{ci}_bmax"] = bmax
        cv[f"c{ci}_voxel_brick"] = np.array([voxel, brick], np.float32)
        cv[f"c{ci}_sel"] = sel.astype(np.uint32)
        cv[f"c{ci}_counts"] = np.array([len(l) for l in lists], np.uint32)
        cv[f"c{ci}_indices"] = np.concatenate(lists).astype(np.uint32)
    cv["positions_7_5_3"] = R.voxel_positions((7, 5, 3))
    rq = np.concatenate([np.linspace(0, 40, 161), rng.uniform(0, 60, 200)]).astype(np.float32)
    cv["round_in"] = rq
    cv["round_out"] = np.array([R.glm_round(v) for v in rq], np.float32)
    np.savez_compressed(os.path.join(OUT, "ref_bricks.npz"), **cv)
    # --- draw uniforms (real gloost::Matrix / glm arithmetic, recon_integration.cpp:183-206)
    views = []
    for eye in ((1.6, 1.5, 2.2), (-2.0, 1.0, 1.2), (0.7, 1.3, 0.75)):
        mv, pr = synth.look_at(eye, (0.0, 1.1, 0.0)), synth.perspective(50.0, 16 / 9, 0.1, 10.0)
        u = R.draw_uniforms(mv, pr, sc.bbox_min, sc.bbox_max, 320, 180)
        views.append(np.concatenate([mv, pr, u["img_to_eye"], u["normal_matrix"], u["camera_pos"]]))
    np.savez_compressed(os.path.join(OUT, "ref_draw_uniforms.npz"), views=np.array(views, np.float32), bbox_min=sc.bbox_min, bbox_max=sc.bbox_max)
    # --- getTrilinear (real DataTypes.cpp) at interior voxel-unit coordinates
    q = rng.uniform(0.0, 1.0, size=(256, 3)) * (np.array([16, 16, 32]) - 1.001)
    tri = np.array([R.get_trilinear(xyz, *map(float, p)) for p in q.astype(np.float32)], np.float32)
    np.savez_compressed(os.path.join(OUT, "ref_trilinear.npz"), coords=q.astype(np.float32), values=tri)
    # --- self-regression pins of the shader restatement (one small fused frame + one raymarch)
    sc2 = synth.make_scene(N=2, W=128, H=106, CW=160, CH=135, cv_res=(32, 32, 64))
    inv2 = synth.analytic_inverse(sc2, (50, 55, 50))
    grid = O.brick_grid(sc2.bbox_min, sc2.bbox_max, 0.02, 0.1)
    cams = [O.frustum(sc2.cv_xyz[i])[1] for i in range(sc2.N)]
    pre = O.preprocess(sc2, grid, cams)
    occ = O.occupied_bricks(pre["bricks"], 10)
    tsdf = O.integrate(inv2, pre, grid, 0.01, True, occ)
    mv, pr = synth.look_at((1.6, 1.5, 2.2), (0.0, 1.1, 0.0)), synth.perspective(50.0, 16 / 9, 0.1, 10.0)
    rm = O.raymarch(tsdf, 0.01, inv2, sc2, pre, grid, occ, mv, pr, 160, 90, 1, True)
    pins = {k: sha(v) for k, v in pre.items()}
    pins.update(occupied=sha(occ), tsdf=sha(tsdf), rgba=sha(rm["rgba"]), depth=sha(rm["depth"]), samples=sha(rm["samples"]))
    np.savez_compressed(os.path.join(OUT, "oracle_frame_pins.npz"), names=np.array(list(pins.keys())), sha256=np.array(list(pins.values())),
                        n_occupied=len(occ), band=int(((tsdf > -0.01) & (tsdf < 0.01)).sum()), hits=int((rm["depth"] < 1).sum()))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
