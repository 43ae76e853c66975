"""Seeded synthetic inputs for the volumetric-fusion path (nothing ships with the reference).

Sensors: N pinhole depth cameras on a ring (radius 2.0 m, height 1.1 m, azimuth 2*pi*i/N + 0.1) looking at the
bounding-box centre; Kinect-v2-like intrinsics scaled to the requested resolution. Forward calibration volumes
`cv_xyz` / `cv_uv` follow the reference layout (framework/calibration/calibration_volume.hpp:18-27: x fastest,
index z*X*Y + y*X + x) with a smooth seeded distortion so nearest-neighbour sets are unique. Scene: analytic SDF
(sphere + capsule torso), sphere-traced per sensor into depth maps in metres (0 = no return), plus Gaussian
noise and random drop-outs; colour is a procedural checker (RGB8).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

DEPTH_LIMITS = (0.5, 4.5)


@dataclasses.dataclass
class Sensor:
    pos: np.ndarray        # camera centre (3,)
    right: np.ndarray      # +u direction
    down: np.ndarray       # +v direction
    fwd: np.ndarray        # viewing direction
    fx: float
    fy: float
    cx: float
    cy: float
    W: int
    H: int
    # colour camera
    cpos: np.ndarray
    cf: float
    ccx: float
    ccy: float
    CW: int
    CH: int


@dataclasses.dataclass
class Scene:
    N: int
    W: int
    H: int
    CW: int
    CH: int
    bbox_min: np.ndarray
    bbox_max: np.ndarray
    sensors: list
    cv_res: tuple          # (X, Y, Z) of the forward volumes
    cv_xyz: np.ndarray     # [N][Z][Y][X][3] float32
    cv_uv: np.ndarray      # [N][Z][Y][X][2] float32
    depth: np.ndarray      # [N][H][W] float32 metres
    color: np.ndarray      # [N][CH][CW][3] uint8


def _normalize(v):
    return v / np.linalg.norm(v)


def make_sensors(N, W, H, CW, CH, bbox_min, bbox_max, radius=2.0, height=1.1):
    centre = 0.5 * (np.asarray(bbox_min, np.float64) + np.asarray(bbox_max, np.float64))
    target = np.array([centre[0], height, centre[2]])
    up = np.array([0.0, 1.0, 0.0])
    out = []
    for i in range(N):
        az = 2.0 * math.pi * i / N + 0.1
        pos = np.array([centre[0] + radius * math.cos(az), height, centre[2] + radius * math.sin(az)])
        fwd = _normalize(target - pos)
        right = _normalize(np.cross(fwd, up))
        down = -np.cross(right, fwd)
        out.append(Sensor(pos=pos, right=right, down=down, fwd=fwd,
                          fx=365.5 * W / 512.0, fy=365.5 * H / 424.0, cx=W / 2.0, cy=H / 2.0, W=W, H=H,
                          cpos=pos + 0.052 * right, cf=1060.0 * CW / 1280.0, ccx=CW / 2.0, ccy=CH / 2.0, CW=CW, CH=CH))
    return out


def _distortion(s, t, r, seed):
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * math.pi, size=(3, 3))
    fr = rng.uniform(1.5, 3.5, size=(3, 3))
    amp = 0.002
    d = np.empty(s.shape + (3,), np.float64)
    for c in range(3):
        d[..., c] = amp * (np.sin(fr[c, 0] * 2 * math.pi * s + ph[c, 0]) + np.sin(fr[c, 1] * 2 * math.pi * t + ph[c, 1])
                           + np.sin(fr[c, 2] * 2 * math.pi * r + ph[c, 2])) / 3.0
    return d


def forward_volumes(sen: Sensor, cv_res, seed):
    X, Y, Z = cv_res
    s = (np.arange(X) + 0.5) / X
    t = (np.arange(Y) + 0.5) / Y
    r = (np.arange(Z) + 0.5) / Z
    R, T, S = np.meshgrid(r, t, s, indexing="ij")            # [Z][Y][X]
    px, py = S * sen.W, T * sen.H
    z = DEPTH_LIMITS[0] + (DEPTH_LIMITS[1] - DEPTH_LIMITS[0]) * R
    ray = (sen.fwd[None, None, None, :] + ((px - sen.cx) / sen.fx)[..., None] * sen.right
           + ((py - sen.cy) / sen.fy)[..., None] * sen.down)
    world = sen.pos + z[..., None] * ray + _distortion(S, T, R, seed)
    rel = world - sen.cpos
    zc = rel @ sen.fwd
    uc = sen.cf * (rel @ sen.right) / zc + sen.ccx
    vc = sen.cf * (rel @ sen.down) / zc + sen.ccy
    uv = np.stack([uc / sen.CW, vc / sen.CH], axis=-1)
    return world.astype(np.float32), uv.astype(np.float32)


def scene_sdf(p, frame=0, centre=(0.0, 1.1, 0.0)):
    """Signed distance (metres) of world points p[..., 3]; frame t rotates the scene by t degrees about +y."""
    if frame:
        a = math.radians(frame)
        c, s = math.cos(a), math.sin(a)
        q = p - np.asarray(centre)
        p = np.stack([c * q[..., 0] + s * q[..., 2], q[..., 1], -s * q[..., 0] + c * q[..., 2]], -1) + np.asarray(centre)
    sph = np.linalg.norm(p - np.array([0.05, 1.45, 0.0]), axis=-1) - 0.28
    a_, b_ = np.array([0.0, 0.45, 0.05]), np.array([0.0, 1.0, 0.0])
    pa, ba = p - a_, b_ - a_
    h = np.clip((pa @ ba) / (ba @ ba), 0.0, 1.0)
    cap = np.linalg.norm(pa - h[..., None] * ba, axis=-1) - 0.26
    arm_a, arm_b = np.array([-0.55, 1.05, 0.0]), np.array([0.55, 0.95, 0.1])
    pa2, ba2 = p - arm_a, arm_b - arm_a
    h2 = np.clip((pa2 @ ba2) / (ba2 @ ba2), 0.0, 1.0)
    arm = np.linalg.norm(pa2 - h2[..., None] * ba2, axis=-1) - 0.09
    return np.minimum(np.minimum(sph, cap), arm)


def wall_sdf(p, frame=0, centre=(0.0, 1.1, 0.0)):
    """Half-space whose boundary plane passes through the bbox centre, perpendicular to sensor 0's optical axis
    (azimuth 0.1): sensor 0 sees a wall at exactly z = 2.0 m in every pixel (analytic checks)."""
    n = np.array([math.cos(0.1), 0.0, math.sin(0.1)])
    return (p - np.asarray(centre)) @ n      # positive on sensor 0's side, negative (inside) behind the plane


def render_depth(sen: Sensor, sdf, frame, seed, noise_sigma=0.0015, dropout=0.01):
    """Sphere-traced z-depth map. Only rays still marching are evaluated each step (same values as marching every pixel
    every step - the SDF is evaluated per point - at a tenth of the time)."""
    W, H = sen.W, sen.H
    px, py = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    ray = (sen.fwd[None, None, :] + ((px - sen.cx) / sen.fx)[..., None] * sen.right + ((py - sen.cy) / sen.fy)[..., None] * sen.down)
    rl = np.linalg.norm(ray, axis=-1)
    d = (ray / rl[..., None]).reshape(-1, 3)
    t = np.full(H * W, 0.3)
    hit = np.zeros(H * W, bool)
    idx = np.arange(H * W)
    for _ in range(96):
        ti = t[idx]
        dist = sdf(sen.pos + ti[:, None] * d[idx], frame)
        h = dist < 2e-4
        hit[idx[h]] = True
        keep = ~h & (ti < 6.0)
        idx = idx[keep]
        t[idx] = ti[keep] + np.maximum(dist[keep], 1e-4)
        if idx.size == 0:
            break
    z = t.reshape(H, W) / rl                                # z-depth along the optical axis
    rng = np.random.default_rng(seed)
    z = z + rng.normal(0.0, noise_sigma, size=z.shape)
    drop = rng.uniform(size=z.shape) < dropout
    z = np.where(hit.reshape(H, W) & ~drop, z, 0.0)
    return z.astype(np.float32)


def render_color(sen: Sensor, seed):
    v, u = np.meshgrid(np.arange(sen.CH), np.arange(sen.CW), indexing="ij")
    sz = max(8, sen.CW // 40)
    chk = (((u // sz) + (v // sz)) % 2).astype(np.float64)
    rng = np.random.default_rng(seed)
    base = rng.uniform(40, 215, size=3)
    img = np.empty((sen.CH, sen.CW, 3), np.float64)
    img[..., 0] = base[0] + 35.0 * chk + 20.0 * u / sen.CW
    img[..., 1] = base[1] - 30.0 * chk + 25.0 * v / sen.CH
    img[..., 2] = base[2] + 15.0 * chk * (u % 7 == 0)
    return np.clip(img, 0, 255).astype(np.uint8)


def make_scene(N=4, W=512, H=424, CW=1280, CH=1080, cv_res=(128, 128, 256), bbox=((-1.0, 0.0, -1.0), (1.0, 2.2, 1.0)),
               frame=0, seed=1234, sdf=scene_sdf) -> Scene:
    bmin, bmax = np.asarray(bbox[0], np.float32), np.asarray(bbox[1], np.float32)
    sensors = make_sensors(N, W, H, CW, CH, bmin, bmax)
    X, Y, Z = cv_res
    cv_xyz = np.empty((N, Z, Y, X, 3), np.float32)
    cv_uv = np.empty((N, Z, Y, X, 2), np.float32)
    depth = np.empty((N, H, W), np.float32)
    color = np.empty((N, CH, CW, 3), np.uint8)
    for i, s in enumerate(sensors):
        cv_xyz[i], cv_uv[i] = forward_volumes(s, cv_res, seed + 100 + i)
        depth[i] = render_depth(s, sdf, frame, seed + i)
        color[i] = render_color(s, seed + 50 + i)
    return Scene(N, W, H, CW, CH, bmin, bmax, sensors, tuple(cv_res), cv_xyz, cv_uv, depth, color)


def rerender(scene: Scene, frame, seed=1234, sdf=scene_sdf) -> Scene:
    """Same rig and calibration, new depth maps for scene frame `frame` (colour unchanged)."""
    depth = np.stack([render_depth(s, sdf, frame, seed + i) for i, s in enumerate(scene.sensors)])
    return dataclasses.replace(scene, depth=depth)


def analytic_inverse(scene: Scene, res) -> np.ndarray:
    """Closed-form stand-in for calib_inverter output: [N][Z][Y][X][4] float32, (u, v, d, 1) in normalised texture
    coordinates of the forward volume for voxel centres inside the sensor frustum, all -1 outside (the convention of
    calibration_inverter.cpp:127-141). Ignores the 2 mm distortion; used where a fast, valid input is enough."""
    X, Y, Z = res
    dims = (scene.bbox_max - scene.bbox_min).astype(np.float64)
    xs = scene.bbox_min[0] + (np.arange(X) + 0.5) / X * dims[0]
    ys = scene.bbox_min[1] + (np.arange(Y) + 0.5) / Y * dims[1]
    zs = scene.bbox_min[2] + (np.arange(Z) + 0.5) / Z * dims[2]
    ZZ, YY, XX = np.meshgrid(zs, ys, xs, indexing="ij")
    P = np.stack([XX, YY, ZZ], -1)
    out = np.empty((scene.N, Z, Y, X, 4), np.float32)
    for i, s in enumerate(scene.sensors):
        rel = P - s.pos
        zc = rel @ s.fwd
        with np.errstate(divide="ignore", invalid="ignore"):
            u = (s.fx * (rel @ s.right) / zc + s.cx) / s.W
            v = (s.fy * (rel @ s.down) / zc + s.cy) / s.H
        d = (zc - DEPTH_LIMITS[0]) / (DEPTH_LIMITS[1] - DEPTH_LIMITS[0])
        ok = (zc > DEPTH_LIMITS[0]) & (zc < DEPTH_LIMITS[1]) & (u > 0) & (u < 1) & (v > 0) & (v < 1)
        o = np.stack([u, v, d, np.ones_like(u)], -1)
        o[~ok] = -1.0
        out[i] = o.astype(np.float32)
    return out


def look_at(eye, target, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """Column-major float32[16] modelview, as glGetFloatv(GL_MODELVIEW_MATRIX) would return it (gluLookAt)."""
    eye, target, up = (np.asarray(v, np.float64) for v in (eye, target, up))
    f = _normalize(target - eye)
    s = _normalize(np.cross(f, up))
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[:3, 3] = -m[:3, :3] @ eye
    return np.ascontiguousarray(m.T, np.float32).reshape(16)


def perspective(fovy_deg, aspect, near, far) -> np.ndarray:
    """Column-major float32[16] projection (gluPerspective)."""
    f = 1.0 / math.tan(math.radians(fovy_deg) / 2.0)
    m = np.zeros((4, 4))
    m[0, 0], m[1, 1] = f / aspect, f
    m[2, 2], m[2, 3] = (far + near) / (near - far), 2.0 * far * near / (near - far)
    m[3, 2] = -1.0
    return np.ascontiguousarray(m.T, np.float32).reshape(16)


# ---- compressed stream formats: what a sender with compress_rgb / compress_depth puts on the wire ----
def encode_depth8(depth_m, near, far):
    """Inverse of pre_depth.fs uncompress() (:51-61): byte = round(255 * sqrt((d - near) / (far - near) - 0.15 * s)),
    s = (far - near) / 255 / ... as the shader defines scaled_near; 0 metres (no measurement) -> byte 0."""
    d = np.asarray(depth_m, np.float64)
    scale = float(far) - float(near)
    sn = scale / 255.0
    t = (d - near) / scale - 0.15 * sn
    dc = np.sqrt(np.clip(t, 0.0, 1.0))
    b = np.clip(np.rint(dc * 255.0), 0, 255)
    b[d <= 0.0] = 0
    return b.astype(np.uint8)


def encode_dxt5(rgb, seed=0):
    """DXT5 (BC3) stream blocks: 8 alpha bytes (arbitrary here: the path never samples alpha) + the colour block of
    encode_dxt1. In a BC3 block the colour half is always read in four-colour mode, whatever the endpoint order.
    uint8 [H][W][3] -> block bytes (16 per 4x4 block)."""
    c = encode_dxt1(rgb).reshape(-1, 8)
    a = np.random.default_rng(seed).integers(0, 256, size=c.shape, dtype=np.uint8)
    return np.concatenate([a, c], axis=1).reshape(-1)


def encode_dxt1(rgb):
    """Minimal DXT1 (BC1) encoder for synthetic streams: per 4x4 block the endpoints are the corners of the colour
    bounding box in 5:6:5, four-colour mode, nearest palette entry per texel. uint8 [H][W][3] -> block bytes."""
    rgb = np.ascontiguousarray(rgb, np.uint8)
    H, W, _ = rgb.shape
    assert H % 4 == 0 and W % 4 == 0
    blk = rgb.reshape(H // 4, 4, W // 4, 4, 3).transpose(0, 2, 1, 3, 4).reshape(-1, 16, 3).astype(np.int32)
    hi, lo = blk.max(axis=1), blk.min(axis=1)

    def to565(c):
        return ((c[:, 0] >> 3) << 11) | ((c[:, 1] >> 2) << 5) | (c[:, 2] >> 3)

    c0, c1 = to565(hi), to565(lo)
    swap = c0 < c1
    c0, c1 = np.where(swap, c1, c0), np.where(swap, c0, c1)
    same = c0 == c1                     # a flat block: three-colour mode, every index 0

    def from565(v):
        r, g, b = (v >> 11) & 31, (v >> 5) & 63, v & 31
        return np.stack([(r << 3) | (r >> 2), (g << 2) | (g >> 4), (b << 3) | (b >> 2)], axis=1)

    e0, e1 = from565(c0), from565(c1)
    pal = np.stack([e0, e1, (2 * e0 + e1) // 3, (e0 + 2 * e1) // 3], axis=1)          # [nb][4][3]
    dist = ((blk[:, :, None, :] - pal[:, None, :, :]) ** 2).sum(axis=3)                # [nb][16][4]
    idx = dist.argmin(axis=2).astype(np.uint32)
    idx[same] = 0
    bits = (idx << (2 * np.arange(16, dtype=np.uint32))[None, :]).sum(axis=1).astype(np.uint32)
    out = np.zeros((blk.shape[0], 8), np.uint8)
    out[:, 0], out[:, 1] = c0 & 255, c0 >> 8
    out[:, 2], out[:, 3] = c1 & 255, c1 >> 8
    for k in range(4):
        out[:, 4 + k] = (bits >> (8 * k)) & 255
    return out.reshape(-1)
