"""ORACLE binding (test infrastructure, NOT product code): ctypes wrapper over oracle/_ref/libref_harness.so, i.e. REAL
reference sources (frustum.cpp, calibration_inverter.cpp, nearest_neighbour_search.cpp, volume_sampler.cpp,
DataTypes.cpp, calibration_volume.hpp, gloost Matrix, glm 0.9.5.3) compiled where they lie under /root/reference by
oracle/Makefile (see oracle/ref_harness.cpp for what is stubbed). Used to pin the restatement in oracle/*.cpp and to
generate tests/golden/*.npz (tools/make_golden.py). available() is False where the library was never built."""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_harness.so")
_LIB = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def available():
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        L.ref_frustum.argtypes = [f32p, f32p, f32p]
        L.ref_frustum_inside.argtypes = [f32p, f32p, C.c_int, i32p]
        L.ref_calib_invert.argtypes = [C.c_char_p, f32p, C.c_uint, C.c_uint, C.c_uint, f32p, f32p, u32p, f32p]
        L.ref_volume_roundtrip.argtypes = [C.c_char_p, C.c_int, u32p, f32p, f32p, f32p, u32p, f32p]
        L.ref_contained_voxels.argtypes = [u32p, f32p, f32p, u32p, C.c_uint]
        L.ref_contained_voxels.restype = C.c_uint
        L.ref_voxel_positions.argtypes = [u32p, f32p]
        L.ref_get_trilinear.argtypes = [f32p, C.c_uint, C.c_uint, C.c_uint, C.c_float, C.c_float, C.c_float, f32p]
        L.ref_glm_round.argtypes = [C.c_float]
        L.ref_glm_round.restype = C.c_float
        L.ref_glm_distance.argtypes = [f32p, f32p]
        L.ref_glm_distance.restype = C.c_float
        L.ref_draw_uniforms.argtypes = [f32p, f32p, f32p, f32p, C.c_uint, C.c_uint, f32p]
        u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        L.ref_squish_compress_dxt1.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.ref_squish_decompress_dxt1.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.ref_squish_storage_dxt1.argtypes = [C.c_int, C.c_int]
        L.ref_squish_storage_dxt1.restype = C.c_int
        if hasattr(L, "ref_squish_storage_dxt5"):
            L.ref_squish_compress_dxt5.argtypes = [u8p, C.c_int, C.c_int, u8p]
            L.ref_squish_decompress_dxt5.argtypes = [u8p, C.c_int, C.c_int, u8p]
            L.ref_squish_storage_dxt5.argtypes = [C.c_int, C.c_int]
            L.ref_squish_storage_dxt5.restype = C.c_int
        _LIB = L
    return _LIB


def corners(cv_xyz_one):
    """getCornerPoints (calibration_inverter.cpp:157-172 / CalibVolumes.cpp:98-113)."""
    Z, Y, X, _ = cv_xyz_one.shape
    ex, ey, ez = X - 1, Y - 1, Z - 1
    idx = [(0, 0, 0), (0, ey, 0), (ex, ey, 0), (ex, 0, 0), (0, 0, ez), (0, ey, ez), (ex, ey, ez), (ex, 0, ez)]
    return np.ascontiguousarray(np.stack([cv_xyz_one[z, y, x] for x, y, z in idx]), np.float32)


def frustum(cv_xyz_one):
    planes, cam = np.zeros((6, 4), np.float32), np.zeros(3, np.float32)
    lib().ref_frustum(corners(cv_xyz_one), planes, cam)
    return planes, cam


def frustum_inside(cv_xyz_one, points):
    pts = np.ascontiguousarray(points, np.float32)
    out = np.zeros(len(pts), np.int32)
    lib().ref_frustum_inside(corners(cv_xyz_one), pts, len(pts), out)
    return out


def calib_invert(cv_xyz_one, bbox_min, bbox_max, out_res):
    Z, Y, X, _ = cv_xyz_one.shape
    ox, oy, oz = [int(v) for v in out_res]
    out = np.zeros((oz, oy, ox, 4), np.float32)
    with tempfile.TemporaryDirectory() as d:
        rc = lib().ref_calib_invert((d + "/").encode(), np.ascontiguousarray(cv_xyz_one, np.float32), X, Y, Z,
                                    np.ascontiguousarray(bbox_min, np.float32), np.ascontiguousarray(bbox_max, np.float32),
                                    np.array([ox, oy, oz], np.uint32), out)
    if rc != 0:
        raise RuntimeError(f"ref_calib_invert failed: {rc}")
    return out


def volume_roundtrip(data, limits=(0.5, 4.5)):
    Z, Y, X, ch = data.shape
    out = np.zeros_like(data, dtype=np.float32)
    res_out, lim_out = np.zeros(3, np.uint32), np.zeros(2, np.float32)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "vol.bin")
        rc = lib().ref_volume_roundtrip(path.encode(), ch, np.array([X, Y, Z], np.uint32), np.array(limits, np.float32),
                                        np.ascontiguousarray(data, np.float32), out, res_out, lim_out)
        raw = open(path, "rb").read()
    if rc != 0:
        raise RuntimeError("ref_volume_roundtrip failed")
    return out, res_out, lim_out, raw


def contained_voxels(dims, pos, size):
    dims = np.array(dims, np.uint32)
    cap = int(dims.prod())
    out = np.zeros(cap, np.uint32)
    n = lib().ref_contained_voxels(dims, np.array(pos, np.float32), np.array(size, np.float32), out, cap)
    return out[:min(n, cap)].copy(), int(n)


def voxel_positions(dims):
    dims = np.array(dims, np.uint32)
    out = np.zeros((int(dims[2]), int(dims[1]), int(dims[0]), 3), np.float32)
    lib().ref_voxel_positions(dims, out)
    return out


def get_trilinear(cv_xyz_one, x, y, z):
    Z, Y, X, _ = cv_xyz_one.shape
    out = np.zeros(3, np.float32)
    lib().ref_get_trilinear(np.ascontiguousarray(cv_xyz_one, np.float32), X, Y, Z, x, y, z, out)
    return out


def glm_round(x):
    return lib().ref_glm_round(np.float32(x))


def draw_uniforms(modelview, projection, bbox_min, bbox_max, vw, vh):
    out = np.zeros(35, np.float32)
    lib().ref_draw_uniforms(np.ascontiguousarray(modelview, np.float32).reshape(16), np.ascontiguousarray(projection, np.float32).reshape(16),
                            np.ascontiguousarray(bbox_min, np.float32), np.ascontiguousarray(bbox_max, np.float32), int(vw), int(vh), out)
    return dict(img_to_eye=out[:16].copy(), normal_matrix=out[16:32].copy(), camera_pos=out[32:35].copy())


def squish_compress_dxt1(rgb):
    """external/squish CompressImage (kDxt1 | kColourRangeFit) of uint8 [H][W][3] -> block bytes."""
    H, W, _ = rgb.shape
    rgba = np.concatenate([rgb, np.full((H, W, 1), 255, np.uint8)], axis=2)
    out = np.zeros(lib().ref_squish_storage_dxt1(W, H), np.uint8)
    lib().ref_squish_compress_dxt1(np.ascontiguousarray(rgba), W, H, out)
    return out


def squish_compress_dxt5(rgb, alpha=None):
    """external/squish CompressImage (kDxt5 | kColourRangeFit) of uint8 [H][W][3] (+ optional alpha [H][W]) -> block bytes."""
    H, W, _ = rgb.shape
    a = np.full((H, W, 1), 255, np.uint8) if alpha is None else np.asarray(alpha, np.uint8).reshape(H, W, 1)
    rgba = np.concatenate([rgb, a], axis=2)
    out = np.zeros(lib().ref_squish_storage_dxt5(W, H), np.uint8)
    lib().ref_squish_compress_dxt5(np.ascontiguousarray(rgba), W, H, out)
    return out


def squish_decompress_dxt5(blocks, W, H):
    """external/squish DecompressImage (kDxt5) -> uint8 [H][W][4]."""
    out = np.zeros((H, W, 4), np.uint8)
    lib().ref_squish_decompress_dxt5(np.ascontiguousarray(blocks, np.uint8), W, H, out)
    return out


def squish_decompress_dxt1(blocks, W, H):
    """external/squish DecompressImage (kDxt1) -> uint8 [H][W][4]."""
    out = np.zeros((H, W, 4), np.uint8)
    lib().ref_squish_decompress_dxt1(np.ascontiguousarray(blocks, np.uint8), W, H, out)
    return out
