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
