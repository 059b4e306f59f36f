"""Synthetic KITTI-shaped LiDAR frames (there is no dataset on the box).

Follows SURVEY.md section 8(d): a 64-beam ring model over a +-45 degree forward field of view
hitting a ground plane at z = -1.73 m, vertical back-drops and 10-30 random boxes (car /
pedestrian / cyclist sized), clipped to the stock OpenPCDet KITTI range
x in [0,70.4], y in [-40,40], z in [-3,1]; then the reference's `sample_points` step
(pcdet/datasets/processor/data_processor.py:182-212): more than `num_points` -> keep every far
(> 40 m) point and a random subset of the near ones; fewer -> PAD BY DUPLICATING random points
(so exact duplicates, hence exact FPS ties, are normal); finally shuffle.

Everything is numpy + a seeded Generator, so a frame is a pure function of (seed, num_points).
"""
import numpy as np

KITTI_RANGE = np.array([0.0, -40.0, -3.0, 70.4, 40.0, 1.0], dtype=np.float32)
_CLASS_SIZES = np.array([[3.9, 1.6, 1.56], [0.8, 0.6, 1.73], [1.76, 0.6, 1.73]], dtype=np.float32)


def _ring_scene(rng, n_az=1024, az_deg=(-45.0, 45.0)):
    elev = np.deg2rad(np.linspace(-24.8, 2.0, 64))
    az = np.deg2rad(np.linspace(az_deg[0], az_deg[1], n_az))
    e, a = np.meshgrid(elev, az, indexing="ij")
    e = e + rng.normal(0, 2e-4, e.shape)
    a = a + rng.normal(0, 2e-4, a.shape)
    # range to the ground plane (sensor 1.73 m above it); upward beams never hit it
    with np.errstate(divide="ignore", invalid="ignore"):
        r_ground = np.where(e < -1e-3, 1.73 / np.tan(-e), np.inf)
    # vertical back-drop (walls, vegetation) at a per-azimuth-sector distance
    sector = rng.uniform(15.0, 75.0, 32)
    r_wall = sector[(np.arange(n_az) * 32 // n_az)][None, :] / (np.maximum(np.cos(a), 0.2) if az_deg[1] - az_deg[0] < 180.0 else 1.0)
    r_wall = r_wall + rng.normal(0, 0.15, r_wall.shape)
    r = np.minimum(r_ground, r_wall)
    r = r + rng.normal(0, 0.01, r.shape)
    x = r * np.cos(e) * np.cos(a)
    y = r * np.cos(e) * np.sin(a)
    z = r * np.sin(e)
    pts = np.stack([x, y, z], -1).reshape(-1, 3)
    return pts[np.isfinite(pts).all(1)]


def _boxes(rng):
    nb = int(rng.integers(10, 31))
    out = []
    for _ in range(nb):
        cls = int(rng.integers(0, 3))
        l, w, h = _CLASS_SIZES[cls] * rng.uniform(0.9, 1.1, 3)
        cx, cy = rng.uniform(5.0, 65.0), rng.uniform(-30.0, 30.0)
        yaw = rng.uniform(-np.pi, np.pi)
        dist = np.hypot(cx, cy)
        npts = int(np.clip(12000.0 / (dist * dist) * (l * h), 8, 1500))
        # points on the two faces that look at the sensor
        u = rng.uniform(-0.5, 0.5, (npts, 2))
        face = rng.integers(0, 2, npts)
        lx = np.where(face == 0, -0.5 * l * np.sign(cx), u[:, 0] * l)
        ly = np.where(face == 0, u[:, 0] * w, -0.5 * w * np.sign(cy if cy != 0 else 1.0))
        lz = u[:, 1] * h + (h / 2 - 1.73)
        c, s = np.cos(yaw), np.sin(yaw)
        px = cx + c * lx - s * ly
        py = cy + s * lx + c * ly
        out.append(np.stack([px, py, lz], -1) + rng.normal(0, 0.01, (npts, 3)))
    return np.concatenate(out, 0)


def sample_points(points, num_points, rng):
    """data_processor.py:182-212 semantics on an (n, C) array."""
    n = len(points)
    if num_points == -1 or n == num_points:
        choice = np.arange(n)
    elif num_points < n:
        depth = np.linalg.norm(points[:, 0:3], axis=1)
        near = np.where(depth < 40.0)[0]
        far = np.where(depth >= 40.0)[0]
        if num_points > len(far):
            near_pick = rng.choice(near, num_points - len(far), replace=False)
            choice = np.concatenate([near_pick, far]) if len(far) > 0 else near_pick
        else:
            choice = rng.choice(np.arange(n), num_points, replace=False)
    else:
        extra = rng.choice(np.arange(n), num_points - n, replace=True)  # duplicates
        choice = np.concatenate([np.arange(n), extra])
    rng.shuffle(choice)
    return points[choice]


def kitti_frame(seed, num_points=16384, n_az=None):
    """One frame: (num_points, 4) float32 = x, y, z, intensity.

    n_az (azimuth steps) defaults to a per-frame draw in [200, 300): 64 x n_az rays give
    roughly 14k-22k in-range returns, so some frames are subsampled and some are padded with
    duplicates, like real KITTI front-view crops around the 16384-point budget."""
    rng = np.random.default_rng(seed)
    if n_az is None:
        n_az = int(rng.integers(200, 300))
    pts = np.concatenate([_ring_scene(rng, n_az), _boxes(rng)], 0)
    lo, hi = KITTI_RANGE[:3], KITTI_RANGE[3:]
    keep = ((pts >= lo) & (pts < hi)).all(1)
    pts = pts[keep]
    inten = rng.uniform(0.0, 1.0, (len(pts), 1))
    frame = np.concatenate([pts, inten], 1).astype(np.float32)
    return np.ascontiguousarray(sample_points(frame, num_points, rng))


def kitti_batch(batch, num_points=16384, first_frame=0, n_az=None):
    """(B, num_points, 4) float32, frame i seeded with 1000 + first_frame + i."""
    return np.stack([kitti_frame(1000 + first_frame + i, num_points, n_az) for i in range(batch)], 0)


WAYMO_RANGE = np.array([-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], dtype=np.float32)


def waymo_frame(seed, num_points=163840):
    """BASELINE configs[4] shape: a 360-degree 64-beam sweep (~2600 azimuth steps) clipped to the 150 m
    Waymo range of SURVEY.md section 8(d), 5 channels (x, y, z, intensity, elongation), then the same
    `sample_points` step."""
    rng = np.random.default_rng(seed)
    n_az = int(rng.integers(2500, 3300))
    pts = np.concatenate([_ring_scene(rng, n_az, (-180.0, 180.0)), _boxes(rng)], 0)
    lo, hi = WAYMO_RANGE[:3], WAYMO_RANGE[3:]
    pts = pts[((pts >= lo) & (pts < hi)).all(1)]
    extra = rng.uniform(0.0, 1.0, (len(pts), 2))
    frame = np.concatenate([pts, extra], 1).astype(np.float32)
    return np.ascontiguousarray(sample_points(frame, num_points, rng))


def waymo_batch(batch, num_points=163840, first_frame=0):
    return np.stack([waymo_frame(3000 + first_frame + i, num_points) for i in range(batch)], 0)


def uniform_batch(batch, num_points, first_frame=0, extent=(70.4, 80.0, 4.0)):
    """Uniform-random stress variant, seed 2000 + frame id."""
    out = []
    for i in range(batch):
        rng = np.random.default_rng(2000 + first_frame + i)
        xyz = rng.uniform(0.0, 1.0, (num_points, 3)) * np.array(extent) + np.array([0.0, -extent[1] / 2, -3.0])
        out.append(np.concatenate([xyz, rng.uniform(0, 1, (num_points, 1))], 1).astype(np.float32))
    return np.stack(out, 0)


def to_pcdet_points(frames):
    """(B,N,4) -> pcdet's collated `points` (B*N, 5) = [batch_idx, x, y, z, intensity]
    (pcdet/datasets/dataset.py:237-244)."""
    B, N, C = frames.shape
    bidx = np.repeat(np.arange(B, dtype=np.float32), N)[:, None]
    return np.concatenate([bidx, frames.reshape(B * N, C)], 1)


def random_boxes(n, seed=0, extent=35.0, clusters=None):
    """(n, 7) float32 [x, y, z, dx, dy, dz, heading] in "score order" (row 0 = best): detector-like
    candidates for NMS tests and benches.  With `clusters`, boxes are jittered copies of that many
    objects (what a point-based head emits: many near-duplicate boxes per object); otherwise they
    are spread uniformly over a square of half-width `extent` metres."""
    rng = np.random.default_rng(4000 + seed)
    sizes = np.array([[3.9, 1.6, 1.56], [0.8, 0.6, 1.73], [1.76, 0.6, 1.73]], np.float32)
    b = np.zeros((n, 7), np.float32)
    if clusters:
        cx = rng.uniform(0, 2 * extent, clusters)
        cy = rng.uniform(-extent, extent, clusters)
        ch = rng.uniform(-np.pi, np.pi, clusters)
        ck = rng.integers(0, 3, clusters)
        which = rng.integers(0, clusters, n)
        b[:, 0] = cx[which] + rng.normal(0, 0.3, n)
        b[:, 1] = cy[which] + rng.normal(0, 0.3, n)
        b[:, 3:6] = sizes[ck[which]] * rng.uniform(0.85, 1.15, (n, 3))
        b[:, 6] = ch[which] + rng.normal(0, 0.15, n)
    else:
        b[:, 0] = rng.uniform(0, 2 * extent, n)
        b[:, 1] = rng.uniform(-extent, extent, n)
        b[:, 3:6] = sizes[rng.integers(0, 3, n)] * rng.uniform(0.8, 1.2, (n, 3))
        b[:, 6] = rng.uniform(-np.pi, np.pi, n)
    b[:, 2] = rng.uniform(-1.5, -0.5, n)
    return b.astype(np.float32)
