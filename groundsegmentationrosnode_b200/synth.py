"""Deterministic synthetic LiDAR scans of the sensor shapes BASELINE.json names.

One ray per (row, col) beam of the sensor described by a `GsParams` — altitude
`min_angle + (H-1-row) * d_altitude`, azimuth `col * d_azimuth`, both jittered inside the cell — is
cast into a random scene: a (tilted, gently undulating) ground plane 1.73 m below the sensor, boxes
and vertical cylinders between 3 m and 60 m, and a tall ring wall so that upward beams return too.
Range noise, per-beam dropout and a uniform intensity are added.  Points come out as float32
{x, y, z, intensity} records (the 16-byte pcl::PointXYZI layout the reference's cloud2range node
reads, src/cloud2range_node.cpp:18) in firing order (column by column), optionally shuffled to
exercise the reference's last-writer-wins pixel rule (src/alfa_gs.cpp:291).

Everything is numpy on the host; the same arrays feed the CPU oracle and the GPU library.
"""
import math

import numpy as np

SEED0 = 0x5EED0000
SENSOR_HEIGHT = 1.73


def _dims(p):
    H = p.n_beams
    W = p.sample_freq * 60 // p.rpm
    d_az = 360.0 / W / 180 * math.pi
    d_alt = (p.max_angle - p.min_angle) / (H - 1)
    return H, W, d_az, d_alt


def _cast(dirs, rng, p):
    """Nearest hit distance of unit rays `dirs` (n,3) from the origin in a random scene."""
    dx, dy, dz = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    n = dirs.shape[0]
    t_best = np.full(n, np.inf)

    # ground: plane through (0,0,-h) with a normal tilted by up to 3 degrees, plus a small undulation
    tilt = math.radians(rng.uniform(-3.0, 3.0))
    tilt_dir = rng.uniform(0, 2 * math.pi)
    nrm = np.array([math.sin(tilt) * math.cos(tilt_dir), math.sin(tilt) * math.sin(tilt_dir), math.cos(tilt)])
    denom = dirs @ nrm
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(denom < -1e-9, -SENSOR_HEIGHT * nrm[2] / denom, np.inf)
    amp = rng.uniform(0.0, 0.06)
    kx, ky = rng.uniform(0.05, 0.4, size=2)
    tf = np.where(np.isfinite(t), t, 0.0)
    bump = amp * np.sin(kx * tf * dx) * np.cos(ky * tf * dy)
    t = np.where(np.isfinite(t), tf * (1.0 + bump / SENSOR_HEIGHT), np.inf)
    t_best = np.minimum(t_best, np.where(t > 0, t, np.inf))

    # axis-aligned boxes (cars, walls): slab test
    n_box = int(rng.integers(8, 28))
    inv = 1.0 / np.where(np.abs(dirs) < 1e-12, 1e-12, dirs)
    for _ in range(n_box):
        dist = rng.uniform(3.0, 60.0)
        ang = rng.uniform(0, 2 * math.pi)
        cx, cy = dist * math.cos(ang), dist * math.sin(ang)
        if rng.random() < 0.3:  # wall segment
            sx, sy, sz = rng.uniform(4, 15), rng.uniform(0.2, 0.5), rng.uniform(2, 6)
            if rng.random() < 0.5:
                sx, sy = sy, sx
        else:  # car / crate
            sx, sy, sz = rng.uniform(1.5, 4.5), rng.uniform(1.5, 2.2), rng.uniform(1.2, 2.5)
        lo = np.array([cx - sx / 2, cy - sy / 2, -SENSOR_HEIGHT])
        hi = np.array([cx + sx / 2, cy + sy / 2, -SENSOR_HEIGHT + sz])
        if lo[0] < 0 < hi[0] and lo[1] < 0 < hi[1]:
            continue  # never enclose the sensor
        t0 = lo * inv
        t1 = hi * inv
        tn = np.minimum(t0, t1).max(axis=1)
        tf = np.maximum(t0, t1).min(axis=1)
        hit = (tn <= tf) & (tn > 0.3)
        t_best = np.where(hit & (tn < t_best), tn, t_best)

    # vertical cylinders (poles, trunks)
    n_cyl = int(rng.integers(2, 12))
    a = dx * dx + dy * dy
    for _ in range(n_cyl):
        dist = rng.uniform(3.0, 50.0)
        ang = rng.uniform(0, 2 * math.pi)
        cx, cy, rad, top = dist * math.cos(ang), dist * math.sin(ang), rng.uniform(0.1, 0.6), rng.uniform(2, 8)
        b = -2 * (dx * cx + dy * cy)
        c = cx * cx + cy * cy - rad * rad
        disc = b * b - 4 * a * c
        with np.errstate(invalid="ignore", divide="ignore"):
            tn = (-b - np.sqrt(disc)) / (2 * a)
        z = tn * dz
        hit = (disc > 0) & (tn > 0.3) & (z > -SENSOR_HEIGHT) & (z < -SENSOR_HEIGHT + top)
        t_best = np.where(hit & (tn < t_best), tn, t_best)

    # ring wall ("urban canyon") so that every beam returns inside max_range
    ring = rng.uniform(35.0, min(70.0, 0.7 * p.max_range))
    with np.errstate(divide="ignore", invalid="ignore"):
        tr = ring / np.sqrt(a)
    t_best = np.minimum(t_best, tr)
    return t_best


def make_frame(p, seed=SEED0, dropout=None, jitter=0.3, shuffle=False, noise_sigma=0.02):
    """One synthetic scan for sensor parameters `p` -> float32 array (N, 4) of {x, y, z, intensity}."""
    H, W, d_az, d_alt = _dims(p)
    rng = np.random.Generator(np.random.PCG64(int(seed)))
    if dropout is None:
        dropout = {16: 0.0, 64: 0.08, 128: 0.01}.get(H, 0.05)
    col = np.repeat(np.arange(W), H)  # firing order: column by column
    row = np.tile(np.arange(H), W)
    alt = p.min_angle + (H - 1 - row) * d_alt + rng.uniform(-jitter, jitter, size=H * W) * d_alt
    az = col * d_az + rng.uniform(-jitter, jitter, size=H * W) * d_az
    # keep jittered altitudes inside the sensor's field of view (the reference gates on it, alfa_gs.cpp:269); the
    # outermost beams stay 2 % of a beam step inside the gates, as real returns do not pile up exactly on a gate
    alt = np.clip(alt, p.min_angle + 0.02 * d_alt, p.max_angle - 0.02 * d_alt)
    dirs = np.stack([np.cos(alt) * np.cos(az), np.cos(alt) * np.sin(az), np.sin(alt)], axis=1)
    t = _cast(dirs, rng, p)
    t = t + rng.normal(0.0, noise_sigma, size=t.shape)
    keep = np.isfinite(t) & (t > p.min_range + 0.05) & (t < p.max_range - 0.05)
    keep &= rng.random(H * W) >= dropout
    pts = np.empty((int(keep.sum()), 4), np.float32)
    pts[:, :3] = (dirs[keep] * t[keep, None]).astype(np.float32)
    pts[:, 3] = rng.random(pts.shape[0]).astype(np.float32)
    if shuffle:
        pts = pts[rng.permutation(pts.shape[0])]
    return np.ascontiguousarray(pts)


def make_pool(p, n_distinct, seed0=SEED0, **kw):
    """`n_distinct` frames -> (points (sum N, 4) float32, offsets int64 (n_distinct + 1))."""
    frames = [make_frame(p, seed0 + i, **kw) for i in range(n_distinct)]
    offsets = np.zeros(n_distinct + 1, np.int64)
    offsets[1:] = np.cumsum([f.shape[0] for f in frames])
    return np.ascontiguousarray(np.concatenate(frames, axis=0)), offsets


def tile_pool(points, offsets, n_frames):
    """Repeat a pool of distinct frames cyclically up to `n_frames` (BASELINE config C4: a pool of 64
    distinct 64x2048 frames tiled to 4096)."""
    n_distinct = len(offsets) - 1
    idx = np.arange(n_frames) % n_distinct
    sizes = (offsets[1:] - offsets[:-1])[idx]
    out_off = np.zeros(n_frames + 1, np.int64)
    out_off[1:] = np.cumsum(sizes)
    reps, rem = divmod(n_frames, n_distinct)
    parts = [points] * reps
    if rem:
        parts.append(points[: offsets[rem]])
    return np.ascontiguousarray(np.concatenate(parts, axis=0)), out_off


def make_adversarial_frame(p, seed=SEED0, kind="edges"):
    """Inputs the reference never tests but its code reacts to:
    'edges'      beams exactly on cell boundaries in altitude and azimuth,
    'collisions' several returns per pixel in random order (last-writer-wins, Q1),
    'nonfinite'  a normal scan with NaN / Inf / out-of-gate points mixed in (Q3),
    'sparse'     85 % dropout with whole empty columns (empty seeds Q7, column-1 quirk Q5)."""
    H, W, d_az, d_alt = _dims(p)
    rng = np.random.Generator(np.random.PCG64(int(seed) ^ 0xADE5))
    if kind == "edges":
        n = 20000
        row_edge = rng.integers(0, H, size=n) + rng.choice([0.0, 0.5, -0.5], size=n)
        col_edge = rng.integers(0, W, size=n) + rng.choice([0.0, 0.5, -0.5], size=n)
        alt = np.clip(p.min_angle + row_edge * d_alt, p.min_angle, p.max_angle)
        az = np.mod(col_edge * d_az, 2 * math.pi)
        r = rng.uniform(p.min_range, p.max_range, size=n)
        # a share of ranges exactly on the encoder's quantisation steps and on the gates
        q = rng.integers(0, 65535, size=n) / 65534.0 * (p.max_range - p.min_range) + p.min_range
        r = np.where(rng.random(n) < 0.3, q, r)
        r[:8] = [p.min_range, p.max_range, p.min_range * (1 - 1e-7), p.max_range * (1 + 1e-7), 1.0, 2.0, 50.0, 99.999]
        pts = np.stack([np.cos(alt) * np.cos(az) * r, np.cos(alt) * np.sin(az) * r, np.sin(alt) * r, rng.random(n)], 1)
        return np.ascontiguousarray(pts.astype(np.float32))
    if kind == "collisions":
        a = make_frame(p, seed, dropout=0.0, jitter=0.45)
        b = make_frame(p, seed + 1, dropout=0.3, jitter=0.45)
        c = make_frame(p, seed + 2, dropout=0.6, jitter=0.45)
        pts = np.concatenate([a, b, c], 0)
        return np.ascontiguousarray(pts[rng.permutation(pts.shape[0])])
    if kind == "nonfinite":
        pts = make_frame(p, seed)
        n = pts.shape[0]
        bad = rng.choice(n, size=min(n, 64), replace=False)
        vals = [np.nan, np.inf, -np.inf]
        for k, i in enumerate(bad):
            pts[i, k % 3] = vals[(k // 3) % 3]
        extra = np.array([[0, 0, 0, 0], [1e-3, 0, 0, 0], [500, 0, 0, 0], [0, 0, 5, 0], [0, 0, -5, 0],
                          [-3.0, -0.0, -0.5, 0], [-3.0, 0.0, -0.5, 0], [3.0, -0.0, -0.5, 0]], np.float32)
        return np.ascontiguousarray(np.concatenate([pts, extra], 0))
    if kind == "sparse":
        pts = make_frame(p, seed, dropout=0.85)
        az = np.mod(np.arctan2(pts[:, 1], pts[:, 0]), 2 * math.pi)
        col = ((az + d_az / 2) / d_az).astype(np.int64)
        empty_cols = np.concatenate([[0], rng.choice(np.arange(2, W), size=W // 16, replace=False)])
        keep = ~np.isin(col % W, empty_cols) & (col < W)
        return np.ascontiguousarray(pts[keep])
    raise ValueError(kind)
