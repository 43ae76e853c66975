"""Calibration-volume files in the reference's binary layout (framework/calibration/calibration_volume.hpp:18-27,50-67):
uint32 res.x, res.y, res.z; float32 depth_min, depth_max; T data[res.x*res.y*res.z] with index z*X*Y + y*X + x.
T = kinect::xyz (3 floats, *.cv_xyz), kinect::uv (2 floats, *.cv_uv), glm::fvec4 (4 floats, *.cv_xyz_inv)."""
import numpy as np


def write_volume(path, data, depth_limits=(0.5, 4.5)):
    data = np.ascontiguousarray(data, np.float32)
    Z, Y, X, _ = data.shape
    with open(path, "wb") as f:
        f.write(np.array([X, Y, Z], np.uint32).tobytes())
        f.write(np.array(depth_limits, np.float32).tobytes())
        f.write(data.tobytes())


def read_volume(path, channels):
    with open(path, "rb") as f:
        res = np.frombuffer(f.read(12), np.uint32)
        lim = np.frombuffer(f.read(8), np.float32)
        n = int(res[0]) * int(res[1]) * int(res[2]) * channels
        data = np.frombuffer(f.read(n * 4), np.float32)
    if data.size != n:
        raise IOError(f"{path}: short read ({data.size} of {n} floats)")
    return data.reshape(int(res[2]), int(res[1]), int(res[0]), channels).copy(), (float(lim[0]), float(lim[1]))
