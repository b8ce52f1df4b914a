"""Deterministic synthetic workload for the M-LOAM hot path.

Workload generator only — used by tests/ and bench.py; it is neither the oracle nor the product.

Scene: axis-aligned room 60 x 40 x 8 m, 40 boxes (1-4 m) standing on the floor, 30 vertical poles
(r = 0.15 m).  Sweeps are analytic ray casts (ring-major, ring id in int(intensity), rel. time in the
fraction — image_segmenter.hpp:128 / feature_extract.cpp:111-112 convention) with N(0, 0.02 m) range
noise; ScanInfo.start = ring_begin + 5, end = ring_end - 6 (image_segmenter.hpp:385-387).

Submaps are sampled directly on the scene geometry instead of being accumulated from 30 ray-cast
keyframes (documented deviation from SURVEY §8d's sketch: it keeps the generator independent of
extractCloud): the surf map is area-uniform on floor/ceiling/walls/box faces, the edge map is
length-uniform on room edges, box edges and 8 surface lines per pole; both get N(0, 1 cm) isotropic
jitter and exactly edge:surf = 1:9 points.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

SEED = 20260923
ROOM_MIN = np.array([-30.0, -20.0, 0.0])
ROOM_MAX = np.array([30.0, 20.0, 8.0])
SCAN_PERIOD = 0.1


@dataclass
class Scene:
    box_min: np.ndarray  # [B,3]
    box_max: np.ndarray  # [B,3]
    pole_xy: np.ndarray  # [P,2]
    pole_r: float


def make_scene(seed: int = SEED, n_boxes: int = 40, n_poles: int = 30) -> Scene:
    rng = np.random.Generator(np.random.PCG64(seed))
    size = rng.uniform(1.0, 4.0, size=(n_boxes, 3))
    cxy = np.stack([rng.uniform(-27, 27, n_boxes), rng.uniform(-17, 17, n_boxes)], axis=1)
    # keep a 3 m corridor along y ~ 0 free for the trajectory
    cxy[:, 1] = np.where(np.abs(cxy[:, 1]) < 3.5, np.sign(cxy[:, 1] + 1e-9) * (3.5 + np.abs(cxy[:, 1])), cxy[:, 1])
    bmin = np.concatenate([cxy - size[:, :2] / 2, np.zeros((n_boxes, 1))], axis=1)
    bmax = np.concatenate([cxy + size[:, :2] / 2, size[:, 2:3]], axis=1)
    pxy = np.stack([rng.uniform(-28, 28, n_poles), rng.uniform(-18, 18, n_poles)], axis=1)
    pxy[:, 1] = np.where(np.abs(pxy[:, 1]) < 2.0, np.sign(pxy[:, 1] + 1e-9) * (2.0 + np.abs(pxy[:, 1])), pxy[:, 1])
    return Scene(bmin, bmax, pxy, 0.15)


# ----------------------------------------------------------------------------- poses
def quat_from_rpy(roll: float, pitch: float, yaw: float) -> np.ndarray:
    cr, sr = math.cos(roll / 2), math.sin(roll / 2)
    cp, sp = math.cos(pitch / 2), math.sin(pitch / 2)
    cy, sy = math.cos(yaw / 2), math.sin(yaw / 2)
    # (x, y, z, w)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy])


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def quat_to_mat(q: np.ndarray) -> np.ndarray:
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def pose7(t, q) -> np.ndarray:
    """Parameter block [tx ty tz qx qy qz qw] (pose_local_parameterization.h:20)."""
    q = np.asarray(q, dtype=np.float64)
    q = q / np.linalg.norm(q)
    return np.concatenate([np.asarray(t, dtype=np.float64), q])


def pose_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    Ra = quat_to_mat(a[3:])
    return pose7(Ra @ b[:3] + a[:3], quat_mul(a[3:], b[3:]))


def pose_inv(a: np.ndarray) -> np.ndarray:
    qi = a[3:] * np.array([-1, -1, -1, 1.0])
    return pose7(-(quat_to_mat(qi) @ a[:3]), qi)


def pose_err(a: np.ndarray, b: np.ndarray) -> tuple[float, float]:
    """(translation error [m], rotation angle [rad]) between two parameter blocks."""
    dt = float(np.linalg.norm(a[:3] - b[:3]))
    qa = a[3:] / np.linalg.norm(a[3:])
    qb = b[3:] / np.linalg.norm(b[3:])
    d = abs(float(np.dot(qa, qb)))
    return dt, 2.0 * math.acos(min(1.0, d))


def trajectory(n_frames: int, seed: int = SEED) -> np.ndarray:
    """Constant twist 1.0 m/s forward, 0.1 rad/s yaw at 10 Hz + N(0, 1 cm / 0.2 deg) jitter."""
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    poses = []
    x, y, yaw = -12.0, 0.0, 0.0
    for _ in range(n_frames):
        jt = rng.normal(0, 0.01, 3)
        jr = rng.normal(0, math.radians(0.2), 3)
        poses.append(pose7([x + jt[0], y + jt[1], 1.8 + jt[2]], quat_from_rpy(jr[0], jr[1], yaw + jr[2])))
        x += 0.1 * math.cos(yaw)
        y += 0.1 * math.sin(yaw)
        yaw += 0.01
    return np.stack(poses)


def perturb_pose(p: np.ndarray, rng: np.random.Generator, sigma_t: float = 0.02, sigma_r_deg: float = 0.3) -> np.ndarray:
    dt = rng.normal(0, sigma_t, 3)
    dr = rng.normal(0, math.radians(sigma_r_deg), 3)
    return pose_mul(p, pose7(dt, quat_from_rpy(*dr)))


# ----------------------------------------------------------------------------- ray casting
def ring_elevations(n_rings: int) -> np.ndarray:
    if n_rings == 16:
        lo, hi = -15.0, 15.0
    elif n_rings == 64:
        lo, hi = -24.8, 2.0
    elif n_rings == 128:
        lo, hi = -25.0, 15.0
    else:
        lo, hi = -20.0, 10.0
    return np.radians(np.linspace(lo, hi, n_rings))


def _cast(scene: Scene, o: np.ndarray, D: np.ndarray) -> np.ndarray:
    """Distance along unit rays D [n,3] from origin o to the first surface (room is closed)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / D
        t_lo = (ROOM_MIN - o) * inv
        t_hi = (ROOM_MAX - o) * inv
        t = np.min(np.maximum(t_lo, t_hi), axis=1)  # exit distance of the enclosing box
        for b in range(scene.box_min.shape[0]):
            t1 = (scene.box_min[b] - o) * inv
            t2 = (scene.box_max[b] - o) * inv
            tn = np.max(np.minimum(t1, t2), axis=1)
            tf = np.min(np.maximum(t1, t2), axis=1)
            hit = (tf >= tn) & (tn > 1e-6)
            t = np.where(hit & (tn < t), tn, t)
        a = D[:, 0] ** 2 + D[:, 1] ** 2
        for p in range(scene.pole_xy.shape[0]):
            oc = o[:2] - scene.pole_xy[p]
            bq = 2.0 * (oc[0] * D[:, 0] + oc[1] * D[:, 1])
            cq = oc[0] ** 2 + oc[1] ** 2 - scene.pole_r ** 2
            disc = bq * bq - 4 * a * cq
            ok = (disc > 0) & (a > 1e-12)
            tp = (-bq - np.sqrt(np.where(ok, disc, 0.0))) / (2 * np.where(ok, a, 1.0))
            z = o[2] + tp * D[:, 2]
            hit = ok & (tp > 1e-6) & (z >= ROOM_MIN[2]) & (z <= ROOM_MAX[2])
            t = np.where(hit & (tp < t), tp, t)
    return t


def make_sweep(scene: Scene, pose: np.ndarray, n_rings: int, horizon: int, seed: int, lidar_id: int = 0,
               ext: np.ndarray | None = None, noise: float = 0.02, max_range: float = 100.0):
    """One LiDAR sweep in the SENSOR frame.

    Returns (cloud float32 [N,4] ring-major, scan_start int32[n_rings], scan_end int32[n_rings]).
    `pose` is base->world; `ext` (optional) is sensor->base.
    """
    T = pose if ext is None else pose_mul(pose, ext)
    R = quat_to_mat(T[3:])
    o = T[:3]
    el = ring_elevations(n_rings)
    az = 2.0 * math.pi * np.arange(horizon) / horizon
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    Ds = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (n_rings, horizon))], axis=2)
    Ds = Ds.reshape(-1, 3)
    Dw = Ds @ R.T
    t = _cast(scene, o, Dw)
    rng = np.random.Generator(np.random.PCG64(seed * 1000 + 7 * lidar_id + 3))
    t = t + rng.normal(0, noise, t.shape)
    keep = (t > 0.5) & (t < max_range)
    ring = np.repeat(np.arange(n_rings), horizon)
    rel = np.tile(np.arange(horizon) / horizon * SCAN_PERIOD * 0.999, n_rings)
    P = (Ds * t[:, None]).astype(np.float32)
    inten = (ring + rel).astype(np.float32)
    cloud = np.concatenate([P, inten[:, None]], axis=1)[keep]
    ring = ring[keep]
    counts = np.bincount(ring, minlength=n_rings)
    begin = np.concatenate([[0], np.cumsum(counts)[:-1]])
    scan_start = (begin + 5).astype(np.int32)
    scan_end = (begin + counts - 6).astype(np.int32)
    return np.ascontiguousarray(cloud, dtype=np.float32), scan_start, scan_end


# ----------------------------------------------------------------------------- submaps
def _faces(scene: Scene):
    """List of (origin, u, v) rectangles: room interior + box tops/sides."""
    F = []
    lo, hi = ROOM_MIN, ROOM_MAX
    ex, ey, ez = np.array([hi[0] - lo[0], 0, 0]), np.array([0, hi[1] - lo[1], 0]), np.array([0, 0, hi[2] - lo[2]])
    F += [(lo, ex, ey), (lo + ez, ex, ey), (lo, ex, ez), (lo + ey, ex, ez), (lo, ey, ez), (lo + ex, ey, ez)]
    for b in range(scene.box_min.shape[0]):
        l, h = scene.box_min[b], scene.box_max[b]
        bx, by, bz = np.array([h[0] - l[0], 0, 0]), np.array([0, h[1] - l[1], 0]), np.array([0, 0, h[2] - l[2]])
        F += [(l + bz, bx, by), (l, bx, bz), (l + by, bx, bz), (l, by, bz), (l + bx, by, bz)]
    return F


def _segments(scene: Scene):
    S = []
    lo, hi = ROOM_MIN, ROOM_MAX

    def box_edges(l, h):
        c = [np.array([x, y, z]) for x in (l[0], h[0]) for y in (l[1], h[1]) for z in (l[2], h[2])]
        E = []
        for i in range(8):
            for j in range(i + 1, 8):
                if np.sum(np.abs(c[i] - c[j]) > 1e-12) == 1:
                    E.append((c[i], c[j]))
        return E

    S += box_edges(lo, hi)
    for b in range(scene.box_min.shape[0]):
        S += box_edges(scene.box_min[b], scene.box_max[b])
    for p in range(scene.pole_xy.shape[0]):
        for k in range(8):
            a = 2 * math.pi * k / 8
            xy = scene.pole_xy[p] + scene.pole_r * np.array([math.cos(a), math.sin(a)])
            S.append((np.array([xy[0], xy[1], lo[2]]), np.array([xy[0], xy[1], hi[2]])))
    return S


def make_submap(scene: Scene, n_total: int, seed: int = SEED, jitter: float = 0.01):
    """Returns (surf_map float32 [Ns,4], corner_map float32 [Nc,4]) with Nc = n_total // 10."""
    rng = np.random.Generator(np.random.PCG64(seed + 2))
    n_edge = n_total // 10
    n_surf = n_total - n_edge
    F = _faces(scene)
    area = np.array([np.linalg.norm(np.cross(u, v)) for _, u, v in F])
    fi = rng.choice(len(F), size=n_surf, p=area / area.sum())
    O = np.stack([f[0] for f in F])[fi]
    U = np.stack([f[1] for f in F])[fi]
    V = np.stack([f[2] for f in F])[fi]
    a, b = rng.random(n_surf), rng.random(n_surf)
    surf = O + a[:, None] * U + b[:, None] * V + rng.normal(0, jitter, (n_surf, 3))
    S = _segments(scene)
    length = np.array([np.linalg.norm(q - p) for p, q in S])
    si = rng.choice(len(S), size=n_edge, p=length / length.sum())
    P0 = np.stack([s[0] for s in S])[si]
    P1 = np.stack([s[1] for s in S])[si]
    c = rng.random(n_edge)
    edge = P0 + c[:, None] * (P1 - P0) + rng.normal(0, jitter, (n_edge, 3))

    def pack(x):
        return np.ascontiguousarray(np.concatenate([x, np.zeros((x.shape[0], 1))], axis=1), dtype=np.float32)

    return pack(surf), pack(edge)


def scan_info_from_cloud(cloud: np.ndarray, n_rings: int):
    ring = cloud[:, 3].astype(np.int32)
    counts = np.bincount(ring, minlength=n_rings)
    begin = np.concatenate([[0], np.cumsum(counts)[:-1]])
    return (begin + 5).astype(np.int32), (begin + counts - 6).astype(np.int32)


# ----------------------------------------------------------------------------- multi-LiDAR rigs
# body_T_laser of estimator/config/config_realvehicle_hercules.yaml:56-59 ("PS-calib", rows are [qx qy qz qw tx ty tz]): the RV rig
RV_EXTRINSICS = np.array([[0, 0, 0, 1, 0, 0, 0],
                          [-0.0169, 0.0575, 0.0195, 0.998, 0.5355, 0.0393, -1.131],
                          [-0.1118, 0.1894, 0.6845, 0.6951, 0.5116, 0.6440, -0.904],
                          [0.0745, 0.1312, -0.7449, 0.6496, 0.4406, -0.628, -1.0295]])


def rig_extrinsics(n_lidars: int) -> np.ndarray:
    """[n_lidars, 7] sensor -> base parameter blocks [t q].  Up to 4 LiDARs: the RV rig (first rows of the RV
    config); more: LiDAR 0 = identity, the others on a 1.2 m ring with +-20 deg tilt."""
    if n_lidars <= 4:
        return np.stack([pose7(r[4:7], r[0:4]) for r in RV_EXTRINSICS[:n_lidars]])
    out = [pose7([0, 0, 0], [0, 0, 0, 1])]
    for k in range(1, n_lidars):
        a = 2 * math.pi * k / n_lidars
        tilt = math.radians(20.0) * (1 if k % 2 else -1)
        out.append(pose7([0.6 * math.cos(a), 0.6 * math.sin(a), 0.0], quat_from_rpy(tilt * math.sin(a), tilt * math.cos(a), a)))
    return np.stack(out)


def make_multi_sweep(scene: Scene, pose: np.ndarray, n_lidars: int, n_rings: int, horizon: int, seed: int, ext: np.ndarray | None = None):
    """The sweeps of all LiDARs of a rig at base pose `pose`, concatenated LiDAR-major.
    Returns (cloud [N,4], scan_start int32[n_lidars*n_rings], scan_end, ext [n_lidars,7]); ScanInfo indexes the concatenation."""
    ext = rig_extrinsics(n_lidars) if ext is None else ext
    clouds, starts, ends, base = [], [], [], 0
    for l in range(n_lidars):
        c, ss, se = make_sweep(scene, pose, n_rings, horizon, seed=seed, lidar_id=l, ext=ext[l])
        clouds.append(c), starts.append(ss + base), ends.append(se + base)
        base += c.shape[0]
    return (np.ascontiguousarray(np.concatenate(clouds)), np.concatenate(starts).astype(np.int32), np.concatenate(ends).astype(np.int32), ext)


# ----------------------------------------------------------------------------- keyframe-built submap
import os

_KF_CACHE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "submap_keyframes_filtered.npz")


def keyframe_map_features(scene: Scene, extract_cloud, voxel_grid, n_keyframes: int = 30, n_rings: int = 64, horizon: int = 2048,
                          use_cache: bool = True):
    """Voxel-filtered edge / surf features of `n_keyframes` earlier keyframes in the world frame: ray-cast sweeps from poses
    1 m apart along the corridor (DISTANCE_KEYFRAMES = 1.0), extractCloud, keyframe pose, pcl::VoxelGrid 0.2 / 0.4
    (MAP_CORNER_RES / MAP_SURF_RES).  `extract_cloud` / `voxel_grid` are passed in (the oracle's: this module stays a pure
    workload generator).  The result for the default arguments is cached in tests/golden/ (made by tests/golden/make_golden.py;
    ~30 s of ray casting otherwise)."""
    tag = f"{SEED}_{n_keyframes}_{n_rings}x{horizon}"
    if use_cache and os.path.exists(_KF_CACHE):
        z = np.load(_KF_CACHE)
        if str(z["tag"]) == tag:
            return z["surf"], z["corner"]
    corner_all, surf_all = [], []
    for k in range(n_keyframes):
        x = -27.0 + 1.0 * k
        kf = pose7([x, 0.3 * math.sin(0.4 * k), 1.8], quat_from_rpy(0.0, 0.0, 0.05 * math.sin(0.3 * k)))
        cloud, ss, se = make_sweep(scene, kf, n_rings, horizon, seed=5000 + k)
        f = extract_cloud(cloud, ss, se)
        R = quat_to_mat(kf[3:])
        for key, acc in (("corner_points_less_sharp", corner_all), ("surf_points_less_flat", surf_all)):
            p = f[key][:, :3].astype(np.float64) @ R.T + kf[:3]
            acc.append(p.astype(np.float32))

    def filt(parts, leaf):
        p = np.concatenate(parts)
        p4 = np.ascontiguousarray(np.concatenate([p, np.zeros((p.shape[0], 1), np.float32)], axis=1), dtype=np.float32)
        return voxel_grid(p4, leaf, True)[0][:, :3].copy()

    return filt(surf_all, 0.4), filt(corner_all, 0.2)


def make_submap_keyframes(scene: Scene, n_total: int, extract_cloud, voxel_grid, seed: int = SEED, jitter: float = 0.01, **kw):
    """The submap as the mapper accumulates it: the voxel-filtered keyframe features re-sampled with N(0, 1 cm)
    jitter to exactly n_total points (edge : surf = 1 : 9).  Returns (surf_map, corner_map, info)."""
    surf_f, corner_f = keyframe_map_features(scene, extract_cloud, voxel_grid, **kw)
    rng = np.random.Generator(np.random.PCG64(seed + 5))

    def resample(ds, n_want):
        m = ds.shape[0]
        if m >= n_want:
            out = ds[np.sort(rng.choice(m, size=n_want, replace=False))].astype(np.float64)
        else:  # up-sample: every filtered point once, then jittered duplicates
            extra = rng.choice(m, size=n_want - m, replace=True)
            out = np.concatenate([ds.astype(np.float64), ds[extra].astype(np.float64) + rng.normal(0, jitter, (n_want - m, 3))])
        return np.ascontiguousarray(np.concatenate([out, np.zeros((n_want, 1))], axis=1), dtype=np.float32)

    n_edge = n_total // 10
    return resample(surf_f, n_total - n_edge), resample(corner_f, n_edge), {"filtered_surf": int(surf_f.shape[0]), "filtered_corner": int(corner_f.shape[0])}


# ----------------------------------------------------------------------------- online-calibration case (config C3)
def transform_cloud(cloud: np.ndarray, T: np.ndarray) -> np.ndarray:
    """T * p for a float32 [n,4] cloud (double math, float store; intensity kept)."""
    out = cloud.copy()
    out[:, :3] = (cloud[:, :3].astype(np.float64) @ quat_to_mat(T[3:]).T + T[:3]).astype(np.float32)
    return out


def make_calib_case(scene: Scene, extract_cloud, voxel_grid, n_rings: int, horizon: int, map_points: int, seed: int = 31, submap=None):
    """Inputs of one online-calibration step (Estimator::optimizeMap, ESTIMATE_EXTRINSIC == 1): the local map in the PIVOT frame, the
    reference LiDAR's features of a later frame i, the second LiDAR's features at the pivot frame (both sensor frame, window-level
    down-sampling 0.2 / 0.4, estimator.cpp:485-496), truth and perturbed initial values (ext_cal off by 2 deg / 5 cm)."""
    traj = trajectory(10)
    pivot, pose_i = traj[3], traj[7]
    ext = rig_extrinsics(2)
    surf_w, corner_w = submap if submap is not None else make_submap(scene, map_points)
    Pinv = pose_inv(pivot)
    surf_map, corner_map = transform_cloud(surf_w, Pinv), transform_cloud(corner_w, Pinv)

    def feats(pose, e, sd):
        cloud, ss, se = make_sweep(scene, pose, n_rings, horizon, seed=sd, lidar_id=0 if e is ext[0] else 1, ext=e)
        f = extract_cloud(cloud, ss, se)
        return voxel_grid(f["surf_points_less_flat"], 0.4, False)[0], voxel_grid(f["corner_points_less_sharp"], 0.2, False)[0]

    surf_ref, corner_ref = feats(pose_i, ext[0], seed)
    surf_cal, corner_cal = feats(pivot, ext[1], seed + 1)
    rng = np.random.Generator(np.random.PCG64(seed))
    pose_i_init = perturb_pose(pose_i, rng)
    d = math.radians(2.0) / math.sqrt(3.0)
    ext_init = pose_mul(ext[1], pose7([0.03, -0.03, 0.027], quat_from_rpy(d, -d, d)))
    return dict(surf_map=surf_map, corner_map=corner_map, surf_ref=surf_ref, corner_ref=corner_ref, surf_cal=surf_cal, corner_cal=corner_cal,
                pivot=pivot, pose_i=pose_i, ext_ref=ext[0], ext_cal=ext[1], pose_i_init=pose_i_init, ext_cal_init=ext_init)
