"""ORACLE binding (test infrastructure, NOT product code): ctypes wrapper over oracle/_ref/libref_glsl.so — the REFERENCE'S
OWN shader sources (glsl/pre_morph.fs, pre_depth.fs, pre_boundary.fs, pre_normal.fs, pre_quality.fs, inc_*.glsl,
tsdf_integration.vs) compiled as C++ against the GLSL host environment oracle/glsl_host/glsl_compat.hpp and run on the CPU
(oracle/glsl_host/glsl_harness.cpp, built by oracle/Makefile where /root/reference is present). Same call shapes as
oracle_py.preprocess / oracle_py.integrate so the restatement and the shaders can be compared on identical inputs.
available() is False where the library was never built."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_glsl.so")
_LIB = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def available():
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        L.rg_pre_morph.argtypes = [f32p, C.c_int, C.c_int, f32p]
        L.rg_pre_depth.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int,
                                   f32p, f32p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, f32p, f32p]
        L.rg_pre_boundary.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, f32p, f32p]
        L.rg_pre_normal.argtypes = [f32p, C.c_int, C.c_int, f32p, C.c_int, C.c_int, C.c_int, f32p, C.c_float, u32p, C.c_uint32, u32p, f32p]
        L.rg_pre_quality.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, f32p, C.c_int, C.c_int, C.c_int, f32p, f32p]
        L.rg_integrate.argtypes = [C.c_int, f32p, i32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_float, u32p, C.c_int, i32p, u32p,
                                   C.c_uint32, f32p]
        _LIB = L
    return _LIB


def preprocess(scene, grid, camera_positions, filter_textures=True, use_processed_depth=True, refine=True, compress=None):
    """NetKinectArray::processDepth + processTextures with the reference's shaders, every layer (see oracle_py.preprocess)."""
    L = lib()
    N, H, W = scene.depth.shape
    X, Y, Z = scene.cv_res
    out = dict(
        morph=np.zeros((N, H, W), np.float32), depth=np.zeros((N, H, W, 2), np.float32),
        lab=np.zeros((N, H, W, 3), np.float32), depth_b=np.zeros((N, H, W, 2), np.float32),
        sil=np.zeros((N, H, W), np.float32), normal=np.zeros((N, H, W, 3), np.float32),
        quality=np.zeros((N, H, W), np.float32), bricks=np.zeros(grid["num_bricks"], np.uint32))
    bmin = np.ascontiguousarray(scene.bbox_min, np.float32)
    bmax = np.ascontiguousarray(scene.bbox_max, np.float32)
    for i in range(N):
        raw = np.ascontiguousarray(scene.depth[i])
        L.rg_pre_morph(raw, W, H, out["morph"][i])
        src = out["morph"][i] if use_processed_depth else raw
        if compress is None:
            cz, scale, near, snear = 0, 0.0, 0.0, 0.0
        else:
            near = np.float32(compress[i][0])
            scale = np.float32(compress[i][1]) - near
            cz, snear = 1, scale / np.float32(255.0)
        L.rg_pre_depth(src, W, H, scene.cv_xyz[i], scene.cv_uv[i], X, Y, Z, scene.color[i], scene.CW, scene.CH, bmin, bmax,
                       0.5, 4.5, int(filter_textures), cz, float(scale), float(near), float(snear), out["depth"][i], out["lab"][i])
        L.rg_pre_boundary(out["depth"][i], out["lab"][i], W, H, int(refine), out["depth_b"][i], out["sil"][i])
        L.rg_pre_normal(out["depth_b"][i], W, H, scene.cv_xyz[i], X, Y, Z, bmin, grid["brick_size"], grid["res_bricks"],
                        grid["num_bricks"], out["bricks"], out["normal"][i])
        L.rg_pre_quality(out["depth_b"][i], out["normal"][i], out["lab"][i], W, H, scene.cv_xyz[i], X, Y, Z,
                         np.ascontiguousarray(camera_positions[i], np.float32), out["quality"][i])
    return out


def integrate(inv, pre, grid, limit, use_bricks, occupied):
    """ReconIntegration::integrate with glsl/tsdf_integration.vs (at most 5 sensors: `uniform sampler3D[5] cv_xyz_inv`)."""
    N, IZ, IY, IX, _ = inv.shape
    assert N <= 5
    _, H, W = pre["sil"].shape
    res = grid["res"]
    tsdf = np.zeros((int(res[2]), int(res[1]), int(res[0])), np.float32)
    occ = np.ascontiguousarray(occupied, np.uint32) if len(occupied) else np.zeros(1, np.uint32)
    lib().rg_integrate(N, np.ascontiguousarray(inv), np.array([IX, IY, IZ], np.int32), np.ascontiguousarray(pre["sil"]),
                       np.ascontiguousarray(pre["depth_b"]), np.ascontiguousarray(pre["quality"]), W, H, np.float32(limit), res,
                       int(use_bricks), grid["ranges"], occ, len(occupied), tsdf)
    return tsdf


def raymarch(tsdf, limit, inv, scene, pre, modelview, projection, width, height, shade_mode=0, depth_peels=None):
    """ReconIntegration::draw with the reference's glsl/tsdf_raymarch.fs + shading.glsl, one fragment per pixel the cube
    proxy covers (skipSpace off unless depth_peels [h][w][4] is given). Host-side uniforms (NormalMatrix, vol_to_world,
    img_to_eye_curr, CameraPos) and the per-pixel ray targets come from the oracle's restatement of
    recon_integration.cpp:177-206, which is pinned separately (tests/golden/ref_draw_uniforms.npz).
    Returns dict(rgba, depth, samples, hit)."""
    import oracle_py as O
    L = lib()
    if not hasattr(L, "_rm"):
        L.rg_raymarch.argtypes = [f32p, u32p, C.c_float, C.c_int, f32p, i32p, f32p, i32p, u8p, C.c_int, C.c_int, f32p, f32p, f32p,
                                  C.c_int, C.c_int, f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, f32p, u8p, C.c_void_p,
                                  f32p, f32p, f32p, u8p]
        L._rm = True
    N, IZ, IY, IX, _ = inv.shape
    X, Y, Z = scene.cv_res
    _, H, W = pre["quality"].shape
    res = np.array([tsdf.shape[2], tsdf.shape[1], tsdf.shape[0]], np.uint32)
    mv = np.ascontiguousarray(modelview, np.float32).reshape(16)
    pr = np.ascontiguousarray(projection, np.float32).reshape(16)
    bmin = np.ascontiguousarray(scene.bbox_min, np.float32)
    bmax = np.ascontiguousarray(scene.bbox_max, np.float32)
    u = O.raymarch_uniforms(mv, pr, bmin, bmax, width, height)
    # gl_NormalMatrix: inverse transpose of the model-view's upper 3x3 (column-major storage), embedded in a mat4
    m3 = mv.reshape(4, 4).T[:3, :3].astype(np.float64)              # row-major 3x3
    n3 = np.linalg.inv(m3).T
    gln = np.eye(4, dtype=np.float64)
    gln[:3, :3] = n3
    gl_normal = np.ascontiguousarray(gln.T.reshape(16), np.float32)  # back to column-major
    v2w = np.zeros(16, np.float32)
    d = (bmax - bmin).astype(np.float32)
    v2w[0], v2w[5], v2w[10], v2w[15] = d[0], d[1], d[2], 1.0
    v2w[12:15] = bmin
    uniforms = np.ascontiguousarray(np.concatenate([mv, pr, gl_normal, u[64:80], v2w, u[0:16]]), np.float32)
    cam = np.ascontiguousarray(u[80:83], np.float32)
    pts, cov = O.raymarch_rays(mv, pr, bmin, bmax, width, height, limit)
    out = dict(rgba=np.zeros((height, width, 4), np.float32), depth=np.zeros((height, width), np.float32),
               samples=np.zeros((height, width), np.float32), hit=np.zeros((height, width), np.uint8), covered=cov)
    peels = np.ascontiguousarray(depth_peels, np.float32) if depth_peels is not None else None
    L.rg_raymarch(np.ascontiguousarray(tsdf), res, np.float32(limit), N, np.ascontiguousarray(inv), np.array([IX, IY, IZ], np.int32),
                  np.ascontiguousarray(scene.cv_uv), np.array([X, Y, Z], np.int32), np.ascontiguousarray(scene.color), scene.CW, scene.CH,
                  np.ascontiguousarray(pre["depth_b"]), np.ascontiguousarray(pre["quality"]), np.ascontiguousarray(pre["normal"]), W, H,
                  bmin, bmax, uniforms, cam, int(width), int(height), int(shade_mode), pts, cov,
                  peels.ctypes.data if peels is not None else None, out["rgba"], out["depth"], out["samples"], out["hit"])
    return out


def fill_colors(rgba, depth, want_atlas=False):
    """ReconIntegration::fillColors with the reference's framebuffer_transfer.fs / tsdf_inpaint.fs / tsdf_colorfill.fs
    (same conventions as oracle_py.fill_colors)."""
    L = lib()
    if not hasattr(L, "_fc"):
        L.rg_fill_colors.argtypes = [f32p, f32p, C.c_int, C.c_int, f32p, C.c_void_p, C.c_void_p]
        L._fc = True
    rgba = np.ascontiguousarray(rgba, np.float32)
    depth = np.ascontiguousarray(depth, np.float32)
    H, W, _ = rgba.shape
    out = np.zeros_like(rgba)
    if not want_atlas:
        L.rg_fill_colors(rgba, depth, W, H, out, None, None)
        return out
    FW = int(np.float32(W) * np.float32(1.5))
    ac = np.zeros((H, FW, 4), np.float32)
    ad = np.zeros((H, FW), np.float32)
    L.rg_fill_colors(rgba, depth, W, H, out, ac.ctypes.data, ad.ctypes.data)
    return out, ac, ad


def depth_peels(scene, grid, occupied, modelview, projection, width, height, limit):
    """What ReconIntegration::drawDepthLimits leaves in the depth-peel texture (recon_integration.cpp:409-429, bricks.vs/fs,
    GL_MIN blending over the clear colour (1, 0, 1, 0)): per pixel (nearest face z, -farthest face z, nearest BACK face z, .)
    in window space, for the full brick_size cube of every occupied brick (bricks.vs:18-19). A rasteriser samples faces at
    pixel centres, which is the intersection of the pixel's ray with the cube; faces in front of the near plane are clipped.
    bricks.gs only drops faces between two occupied bricks, which never hold the per-pixel minimum or maximum. float64."""
    import oracle_py as O
    mv = np.asarray(modelview, np.float64).reshape(4, 4).T              # column-major storage -> row-major matrix
    pr = np.asarray(projection, np.float64).reshape(4, 4).T
    bmin = np.asarray(scene.bbox_min, np.float64)
    dims = np.asarray(scene.bbox_max, np.float64) - bmin
    v2w = np.eye(4)
    v2w[:3, :3] = np.diag(dims)
    v2w[:3, 3] = bmin
    pvm = pr @ mv @ v2w
    u = O.raymarch_uniforms(modelview, projection, scene.bbox_min, scene.bbox_max, width, height)
    cam = u[80:83].astype(np.float64)
    pts, _ = O.raymarch_rays(modelview, projection, scene.bbox_min, scene.bbox_max, width, height, limit)
    d = pts.astype(np.float64) - cam
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rb = [int(v) for v in grid["res_bricks"]]
    bs = float(grid["brick_size"])
    peels = np.zeros((height, width, 4), np.float64)
    peels[..., 0] = 1.0
    peels[..., 2] = 1.0

    def window_z(t):
        p = cam + d * t[..., None]
        clip = p @ pvm[:, :3].T + pvm[:, 3]
        visible = clip[..., 2] >= -clip[..., 3]                          # not clipped by the near plane
        return clip[..., 2] / clip[..., 3] * 0.5 + 0.5, visible

    with np.errstate(divide="ignore", invalid="ignore"):
        inv_d = 1.0 / d
        for bid in np.asarray(occupied, np.int64):
            iz, rem = divmod(int(bid), rb[0] * rb[1])
            iy, ix = divmod(rem, rb[0])
            lo = np.array([ix, iy, iz], np.float64) * bs / dims
            hi = np.array([ix + 1, iy + 1, iz + 1], np.float64) * bs / dims
            ta, tb = (lo - cam) * inv_d, (hi - cam) * inv_d
            t0 = np.minimum(ta, tb).max(axis=-1)
            t1 = np.maximum(ta, tb).min(axis=-1)
            hit = t0 <= t1
            z0, vis0 = window_z(t0)                                      # entry: a front face
            z1, vis1 = window_z(t1)                                      # exit: a back face
            f = hit & vis0 & (t0 > 0)
            b = hit & vis1 & (t1 > 0)
            peels[..., 0] = np.where(f, np.minimum(peels[..., 0], z0), peels[..., 0])
            peels[..., 1] = np.where(f, np.minimum(peels[..., 1], -z0), peels[..., 1])
            peels[..., 0] = np.where(b, np.minimum(peels[..., 0], z1), peels[..., 0])
            peels[..., 1] = np.where(b, np.minimum(peels[..., 1], -z1), peels[..., 1])
            peels[..., 2] = np.where(b, np.minimum(peels[..., 2], z1), peels[..., 2])
    return peels.astype(np.float32)


def reference_cube_strip(path="/root/reference/framework/rendering/unit_cube.cpp"):
    """The reference's unit-cube vertices [8][3] and the triangle strip UnitCube::drawInstanced draws (unit_cube.cpp:21-30,
    53-62), parsed from its source where the tree is present; None otherwise."""
    import re
    if not os.path.exists(path):
        return None
    src = open(path).read()
    verts = re.search(r"std::vector<float> vertices\{(.*?)\};", src, re.S).group(1)
    V = np.array([float(x.rstrip("f")) for x in re.findall(r"[-0-9.]+f", verts)], np.float32).reshape(-1, 3)
    body = src[src.index("void UnitCube::drawInstanced"):]
    idx = np.array([int(x) for x in re.search(r"indices \{(.*?)\};", body, re.S).group(1).replace("\n", " ").split(",")], np.uint8)
    return V, idx


def depth_peels_rasterised(scene, grid, counters, occupied, modelview, projection, width, height, cube=None):
    """ReconIntegration::drawDepthLimits with the reference's bricks.vs / bricks.gs / bricks.fs and a rasteriser
    (glsl_harness.cpp::rg_depth_peels). cube = (vertices [8][3], strip indices); default: parsed from the reference tree."""
    L = lib()
    if not hasattr(L, "_dp"):
        L.rg_depth_peels.argtypes = [f32p, f32p, f32p, C.c_float, u32p, u32p, C.c_uint32, u32p, C.c_uint32, f32p, u8p, C.c_int,
                                     C.c_int, C.c_int, f32p]
        L._dp = True
    cube = cube or reference_cube_strip()
    assert cube is not None, "the unit-cube strip comes from the reference tree"
    V, idx = cube
    out = np.zeros((height, width, 4), np.float32)
    occ = np.ascontiguousarray(occupied, np.uint32) if len(occupied) else np.zeros(1, np.uint32)
    L.rg_depth_peels(np.ascontiguousarray(modelview, np.float32).reshape(16), np.ascontiguousarray(projection, np.float32).reshape(16),
                     np.ascontiguousarray(scene.bbox_min, np.float32), np.float32(grid["brick_size"]), grid["res_bricks"],
                     np.ascontiguousarray(counters, np.uint32), len(counters), occ, len(occupied), np.ascontiguousarray(V, np.float32),
                     np.ascontiguousarray(idx, np.uint8), len(idx), int(width), int(height), out)
    return out


def _gl_normal_matrix(mv):
    """gl_NormalMatrix: inverse transpose of the model-view's upper 3x3 (column-major storage), embedded in a mat4."""
    m3 = mv.reshape(4, 4).T[:3, :3].astype(np.float64)
    gln = np.eye(4, dtype=np.float64)
    gln[:3, :3] = np.linalg.inv(m3).T
    return np.ascontiguousarray(gln.T.reshape(16), np.float32)


def draw_points(scene, pre, modelview, projection, width, height, shade_mode=0):
    """ReconPoints::draw with the reference's glsl/points.vs, points.gs, points.fs (+ shading.glsl, inc_bbox_test.glsl) and the
    fixed-function point pipeline of glsl_harness.cpp::raster_point_gl. Returns (rgba [h,w,4], depth [h,w])."""
    import oracle_py as O
    L = lib()
    if not hasattr(L, "_pts"):
        L.rg_draw_points.argtypes = [C.c_int, C.c_int, C.c_int, f32p, f32p, u8p, C.c_int, C.c_int, f32p, f32p, i32p, f32p, f32p, f32p,
                                     C.c_int, C.c_int, C.c_int, f32p, f32p]
        L.rg_draw_calibs.argtypes = [f32p, u32p, C.c_int, f32p, i32p, f32p, i32p, C.c_int, C.c_float, f32p, f32p, f32p, f32p,
                                     C.c_int, C.c_int, f32p, f32p]
        L._pts = True
    X, Y, Z = scene.cv_res
    N, H, W = pre["quality"].shape
    mv = np.ascontiguousarray(modelview, np.float32).reshape(16)
    pr = np.ascontiguousarray(projection, np.float32).reshape(16)
    bmin, bmax = np.ascontiguousarray(scene.bbox_min, np.float32), np.ascontiguousarray(scene.bbox_max, np.float32)
    u = O.raymarch_uniforms(mv, pr, bmin, bmax, width, height)
    inv4 = lambda m: np.ascontiguousarray(np.linalg.inv(m.reshape(4, 4).T.astype(np.float64)).T.reshape(16), np.float32)     # column-major in and out
    uniforms = np.ascontiguousarray(np.concatenate([mv, pr, _gl_normal_matrix(mv), u[0:16], inv4(pr), inv4(mv)]), np.float32)
    rgba, depth = np.zeros((height, width, 4), np.float32), np.zeros((height, width), np.float32)
    L.rg_draw_points(N, W, H, np.ascontiguousarray(pre["depth_b"], np.float32), np.ascontiguousarray(pre["normal"], np.float32),
                     np.ascontiguousarray(scene.color), scene.CW, scene.CH, np.ascontiguousarray(scene.cv_xyz, np.float32),
                     np.ascontiguousarray(scene.cv_uv, np.float32), np.array([X, Y, Z], np.int32), bmin, bmax, uniforms,
                     int(width), int(height), int(shade_mode), rgba, depth)
    return rgba, depth


def draw_trigrid(scene, pre, modelview, projection, width, height, shade_mode=0, min_length=0.0125):
    """ReconTrigrid::draw with the reference's glsl/trigrid_accum.vs, trigrid_accum.gs, trigrid_accum.fs (+ shading.glsl,
    inc_bbox_test.glsl) and trigrid_normalize.fs through the fixed-function stages of oracle/ro_raster.h.
    Returns (rgba [h,w,4], depth [h,w])."""
    import oracle_py as O
    L = lib()
    L.rg_draw_trigrid.argtypes = [C.c_int, C.c_int, C.c_int, f32p, f32p, u8p, C.c_int, C.c_int, f32p, f32p, i32p, f32p, f32p, f32p,
                                  C.c_int, C.c_int, C.c_int, C.c_float, f32p, f32p]
    X, Y, Z = scene.cv_res
    N, H, W = pre["quality"].shape
    mv = np.ascontiguousarray(modelview, np.float32).reshape(16)
    pr = np.ascontiguousarray(projection, np.float32).reshape(16)
    bmin, bmax = np.ascontiguousarray(scene.bbox_min, np.float32), np.ascontiguousarray(scene.bbox_max, np.float32)
    u = O.raymarch_uniforms(mv, pr, bmin, bmax, width, height)
    uniforms = np.ascontiguousarray(np.concatenate([mv, pr, _gl_normal_matrix(mv), u[0:16]]), np.float32)
    rgba, depth = np.zeros((height, width, 4), np.float32), np.zeros((height, width), np.float32)
    L.rg_draw_trigrid(N, W, H, np.ascontiguousarray(pre["depth_b"], np.float32), np.ascontiguousarray(pre["quality"], np.float32),
                      np.ascontiguousarray(scene.color), scene.CW, scene.CH, np.ascontiguousarray(scene.cv_xyz, np.float32),
                      np.ascontiguousarray(scene.cv_uv, np.float32), np.array([X, Y, Z], np.int32), bmin, bmax, uniforms,
                      int(width), int(height), int(shade_mode), float(min_length), rgba, depth)
    return rgba, depth


def draw_calibs(tsdf, inv, scene, layer, limit, modelview, projection, width, height):
    """ReconCalibs::draw with the reference's glsl/calib_vis.vs and calib_vis.fs over VolumeSampler's voxel centres."""
    draw_points.__doc__
    L = lib()
    if not hasattr(L, "_pts"):
        L.rg_draw_points.argtypes = [C.c_int, C.c_int, C.c_int, f32p, f32p, u8p, C.c_int, C.c_int, f32p, f32p, i32p, f32p, f32p, f32p,
                                     C.c_int, C.c_int, C.c_int, f32p, f32p]
        L.rg_draw_calibs.argtypes = [f32p, u32p, C.c_int, f32p, i32p, f32p, i32p, C.c_int, C.c_float, f32p, f32p, f32p, f32p,
                                     C.c_int, C.c_int, f32p, f32p]
        L._pts = True
    N, IZ, IY, IX, _ = inv.shape
    X, Y, Z = scene.cv_res
    res = np.array([tsdf.shape[2], tsdf.shape[1], tsdf.shape[0]], np.uint32)
    rgba, depth = np.zeros((height, width, 4), np.float32), np.zeros((height, width), np.float32)
    L.rg_draw_calibs(np.ascontiguousarray(tsdf, np.float32), res, N, np.ascontiguousarray(inv, np.float32), np.array([IX, IY, IZ], np.int32),
                     np.ascontiguousarray(scene.cv_xyz, np.float32), np.array([X, Y, Z], np.int32), int(layer), np.float32(limit),
                     np.ascontiguousarray(scene.bbox_min, np.float32), np.ascontiguousarray(scene.bbox_max, np.float32),
                     np.ascontiguousarray(modelview, np.float32).reshape(16), np.ascontiguousarray(projection, np.float32).reshape(16),
                     int(width), int(height), rgba, depth)
    return rgba, depth
