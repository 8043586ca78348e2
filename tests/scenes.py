"""Seeded synthetic inputs: scripted SetOccupancy replays and analytic depth / LIDAR frames.

Everything here is plain numpy and is shared by the parity tests, bench.py and the golden-vector generator; the same
arrays are fed to the CPU oracle and to the CUDA path.
"""
import numpy as np

PARAMS_TOGGLE = (0.97, 0.03, 0.30, 0.90, 0.80)   # one observation toggles a voxel
PARAMS_DEFAULT = (0.70, 0.35, 0.12, 0.97, 0.80)  # parameters.cpp:89-93 / launch files

# parameters.cpp:21-24
FX, FY, CX, CY = 384.458089392, 383.982755697, 322.477357419, 237.076346481


def all_voxels(gs):
    g = np.stack(np.meshgrid(np.arange(gs[0]), np.arange(gs[1]), np.arange(gs[2]), indexing="ij"), -1)
    return g.reshape(-1, 3).astype(np.int32)


def pillar_sites():
    return [(x, y) for x in (12, 22, 32, 42, 52) for y in (12, 22, 32, 42, 52)]


def pillar(x, y):
    return np.array([[x, y, z] for z in range(25)], np.int32)


class Scene:
    """Axis-aligned room (seen from inside) + boxes (seen from outside); `moving` boxes translate every frame."""

    def __init__(self, room_half, n_boxes, n_moving, seed, edge=(0.2, 1.0)):
        rng = np.random.default_rng(seed)
        self.room = np.asarray(room_half, float)
        c = rng.uniform(-self.room * 0.9, self.room * 0.9, (n_boxes, 3))
        e = rng.uniform(edge[0], edge[1], (n_boxes, 3)) / 2
        # keep a free bubble around the origin for the sensor
        keep = np.linalg.norm(c[:, :2], axis=1) > 1.5
        self.c, self.e = c[keep], e[keep]
        self.vel = np.zeros_like(self.c)
        k = min(n_moving, len(self.c))
        ang = rng.uniform(0, 2 * np.pi, k)
        self.vel[:k, 0] = 0.05 * np.cos(ang)
        self.vel[:k, 1] = 0.05 * np.sin(ang)

    def step(self):
        self.c = self.c + self.vel
        out = np.abs(self.c) > self.room[None, :] * 0.9
        self.vel = np.where(out, -self.vel, self.vel)

    def ranges(self, org, dirs):
        """Distance along unit `dirs` (n,3) from `org` to the first surface."""
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / dirs
            t1 = (self.room[None, :] - org[None, :]) * inv
            t2 = (-self.room[None, :] - org[None, :]) * inv
            t = np.where(dirs > 0, t1, np.where(dirs < 0, t2, np.inf)).min(axis=1)
            for c, e in zip(self.c, self.e):
                a = (c - e - org)[None, :] * inv
                b = (c + e - org)[None, :] * inv
                lo = np.nanmax(np.minimum(a, b), axis=1)
                hi = np.nanmin(np.maximum(a, b), axis=1)
                hit = (lo <= hi) & (lo > 0)
                t = np.where(hit & (lo < t), lo, t)
        return t


def pose_walk(n_frames, seed, clamp=2.0, z0=0.0, offset=(0.0137, -0.0211, 0.0093)):
    """Random-walk poses: position step N(0, 0.05 m) clamped to +-clamp, yaw step N(0, 2 deg), pitch = roll = 0.
    `offset` keeps the sensor origin off the voxel lattice planes (the reference DDA never returns for a start exactly on
    a lattice plane heading in a negative direction; see oracle/esdf_oracle.c)."""
    rng = np.random.default_rng(seed)
    p = np.array([offset[0], offset[1], z0 + offset[2]])
    yaw = 0.0
    out = []
    for _ in range(n_frames):
        out.append((p.copy(), yaw))
        p[:2] = np.clip(p[:2] + rng.normal(0, 0.05, 2), -clamp, clamp)
        p[2] = np.clip(p[2] + rng.normal(0, 0.01), z0 - 0.3, z0 + 0.3)
        yaw += np.deg2rad(rng.normal(0, 2.0))
    return out


def rot_z(yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def camera_transform(p, yaw):
    """transform_ (Fiesta.h:415-419) for a camera looking along world +x (rotated by yaw): camera z forward, x right, y down."""
    R = rot_z(yaw) @ np.array([[0, 0, 1.0], [-1, 0, 0], [0, -1, 0]])
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = p
    return T


def body_transform(p, yaw):
    T = np.eye(4)
    T[:3, :3] = rot_z(yaw)
    T[:3, 3] = p
    return T


def depth_frame(scene, p, yaw, width=640, height=480, scale=1.0, max_depth=10.0):
    """Point cloud of Fiesta::DepthConversion without the depth filter (Fiesta.h:341-351): uint16 millimetre depth,
    back-projected with the pin-hole intrinsics (scaled by `scale` for reduced-resolution test images)."""
    fx, fy, cx, cy = FX * scale, FY * scale, CX * scale, CY * scale
    u, v = np.meshgrid(np.arange(width), np.arange(height))
    d_cam = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u, float)], -1).reshape(-1, 3)
    T = camera_transform(p, yaw)
    d_w = d_cam @ T[:3, :3].T
    nrm = np.linalg.norm(d_w, axis=1)
    t = scene.ranges(np.asarray(p, float), d_w / nrm[:, None])
    z = t / nrm                                   # depth along the optical axis
    mm = np.clip(np.round(z * 1000.0), 0, 65535).astype(np.uint16)
    depth = mm.astype(np.float64) / 1000.0
    pts = np.empty((len(depth), 3), np.float32)
    pts[:, 0] = ((u.reshape(-1) - cx) * depth / fx).astype(np.float32)
    pts[:, 1] = ((v.reshape(-1) - cy) * depth / fy).astype(np.float32)
    pts[:, 2] = depth.astype(np.float32)
    return pts, T


def depth_image(scene, p, yaw, width=640, height=480, scale=1.0):
    """uint16 millimetre depth image (sensor_msgs::Image TYPE_16UC1) of the scene + the camera transform_."""
    fx, fy, cx, cy = FX * scale, FY * scale, CX * scale, CY * scale
    u, v = np.meshgrid(np.arange(width), np.arange(height))
    d_cam = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u, float)], -1).reshape(-1, 3)
    T = camera_transform(p, yaw)
    d_w = d_cam @ T[:3, :3].T
    nrm = np.linalg.norm(d_w, axis=1)
    t = scene.ranges(np.asarray(p, float), d_w / nrm[:, None])
    mm = np.clip(np.round(t / nrm * 1000.0), 0, 65535).astype(np.uint16)
    return mm.reshape(height, width), T


def lidar_frame(scene, p, yaw, beams=64, azimuths=1563, noise_seed=None):
    """64 beams, elevation uniformly spaced in [-15, +15] deg, x 1563 azimuths = 100 032 points."""
    el = np.deg2rad(np.linspace(-15.0, 15.0, beams))
    az = 2 * np.pi * np.arange(azimuths) / azimuths
    E, A = np.meshgrid(el, az, indexing="ij")
    d_b = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)
    T = body_transform(p, yaw)
    d_w = d_b @ T[:3, :3].T
    t = scene.ranges(np.asarray(p, float), d_w)
    if noise_seed is not None:
        t = t + np.random.default_rng(noise_seed).normal(0, 0.01, len(t))
    pts = (d_b * t[:, None]).astype(np.float32)
    return pts, T
