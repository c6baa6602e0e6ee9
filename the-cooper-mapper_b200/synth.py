"""Seeded synthetic LiDAR data: procedural scenes, VLP-16 / HDL-64E / tilted-2D scan simulation, map sampling.

There is no network and the reference ships no recorded data, so every test and bench input
comes from here.  Frames are in the sensor's native frame (x forward, y left, z up); the organised entry
(OrganisedScanRegistration::process, OrganizedScanRegistration.cpp:82-150) consumes them as is, the raw-sweep entry
(MultiScanRegistration::process, MultiScanRegistration.cpp:95-200) applies its own (x,y,z) <- (y,z,x) swap.
"""
import numpy as np

LIDARS = {
    # name: (rings, cols, lower_deg, upper_deg, max_range)   MultiScanRegistration.h:90-102
    "VLP-16": (16, 1800, -15.0, 15.0, 100.0),
    "HDL-32": (32, 2048, -30.67, 10.67, 100.0),
    "HDL-64E": (64, 2048, -24.9, 2.0, 120.0),
    "Pandar40": (40, 1800, -15.444, 6.96, 100.0),   # MultiScanMapperP::Pandar40, MultiScanRegistration.h:39-41
}
# beams that are not equally spaced: elevation per ring in degrees, lowest first (Pandar40 data sheet, lidar_type.h:11-52)
ELEVATIONS = {
    "Pandar40": [-15.444, -14.543, -13.63, -12.705, -11.772, -10.826, -9.871, -8.908, -7.934, -6.957, -5.974, -5.647, -5.311, -4.986,
                 -4.657, -4.321, -3.996, -3.663, -3.327, -3.0, -2.667, -2.331, -2.001, -1.667, -1.334, -1.001, -0.667, -0.334, 0.0,
                 0.333, 0.667, 1.001, 1.333, 1.667, 2.001, 2.999, 3.996, 4.988, 5.976, 6.96],
}


class Scene:
    """Ground plane z = ground_z, oriented boxes (buildings) and vertical cylinders (poles) standing on it.

    The ground is NOT at z = 0: LOAM's plane fit solves A n = -1 (feature_utils.h:161-182), which cannot represent a
    plane through the map origin, so the world origin sits at sensor height like a real run that starts at identity.
    """

    def __init__(self, boxes, poles, extent, ground_z=-1.8):
        self.boxes = np.asarray(boxes, np.float64).reshape(-1, 7)   # cx, cy, yaw, hx, hy, z0, z1 (heights above ground)
        self.poles = np.asarray(poles, np.float64).reshape(-1, 4)   # cx, cy, radius, height
        self.extent = float(extent)
        self.ground_z = float(ground_z)
        self.boxes[:, 5] += self.ground_z
        self.boxes[:, 6] += self.ground_z


def make_scene(seed=0, extent=120.0, n_boxes=40, n_poles=30, keep_clear=6.0, corridor=False):
    rng = np.random.default_rng(seed)
    boxes, poles = [], []
    if corridor:   # two long parallel walls: degenerate along x (SURVEY 8c known-answer case)
        boxes.append([0.0, 6.0, 0.0, extent, 0.5, 0.0, 6.0])
        boxes.append([0.0, -6.0, 0.0, extent, 0.5, 0.0, 6.0])
        return Scene(boxes, poles, extent)
    tries = 0
    while len(boxes) < n_boxes and tries < 10000:
        tries += 1
        cx, cy = rng.uniform(-extent, extent, 2)
        hx, hy = rng.uniform(2.0, 12.0, 2)
        if abs(cy) < keep_clear + max(hx, hy) * 1.5:   # keep a street along x free for the trajectory
            continue
        yaw = rng.choice([0.0, rng.uniform(-0.6, 0.6)])
        boxes.append([cx, cy, yaw, hx, hy, 0.0, rng.uniform(3.0, 18.0)])
    while len(poles) < n_poles:
        cx = rng.uniform(-extent, extent)
        cy = rng.choice([-1, 1]) * rng.uniform(3.0, keep_clear + 2.0)
        poles.append([cx, cy, rng.uniform(0.08, 0.25), rng.uniform(3.0, 8.0)])
    return Scene(boxes, poles, extent)


def ray_dirs(model, rows=None, cols=None, tilt_deg=None):
    """Unit ray directions (rows, cols, 3) in the sensor frame; azimuth DEcreases with the column (clockwise spin)."""
    if tilt_deg is None:
        R, Cn, lo, up, _ = LIDARS[model]
        rows = rows or R; cols = cols or Cn
        el = np.deg2rad(np.linspace(lo, up, rows))
        if model in ELEVATIONS and rows == len(ELEVATIONS[model]):
            el = np.deg2rad(np.asarray(ELEVATIONS[model], np.float64))
        az = -2.0 * np.pi * (np.arange(cols) + 0.25) / cols
        ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
        d = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (rows, cols))], -1)
        return d
    # tilted planar scanner (RPLidar on a nodding unit): row = one revolution at a fixed tilt about the y axis
    tilts = np.deg2rad(np.asarray(tilt_deg, np.float64))
    az = -2.0 * np.pi * (np.arange(cols) + 0.25) / cols
    planar = np.stack([np.cos(az), np.sin(az), np.zeros_like(az)], -1)   # (cols, 3)
    out = np.empty((len(tilts), cols, 3))
    for i, t in enumerate(tilts):
        Ry = np.array([[np.cos(t), 0, np.sin(t)], [0, 1, 0], [-np.sin(t), 0, np.cos(t)]])
        out[i] = planar @ Ry.T
    return out


def _cast(scene, o, d, max_range):
    """Ray cast: o (3,), d (N,3) world-frame unit vectors -> range (N,), inf where nothing is hit."""
    N = d.shape[0]
    best = np.full(N, np.inf)
    # ground
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (scene.ground_z - o[2]) / d[:, 2]
    ok = (d[:, 2] < 0) & (t > 0)
    hx = o[0] + t * d[:, 0]; hy = o[1] + t * d[:, 1]
    ok &= (np.abs(hx) <= scene.extent * 1.5) & (np.abs(hy) <= scene.extent * 1.5)
    best = np.where(ok, t, best)
    # boxes (slab test in the box frame)
    for cx, cy, yaw, bx, by, z0, z1 in scene.boxes:
        c, s = np.cos(yaw), np.sin(yaw)
        ox = c * (o[0] - cx) + s * (o[1] - cy); oy = -s * (o[0] - cx) + c * (o[1] - cy); oz = o[2]
        dx = c * d[:, 0] + s * d[:, 1]; dy = -s * d[:, 0] + c * d[:, 1]; dz = d[:, 2]
        tmin = np.full(N, -np.inf); tmax = np.full(N, np.inf)
        for oo, dd, lo, hi in ((ox, dx, -bx, bx), (oy, dy, -by, by), (oz, dz, z0, z1)):
            with np.errstate(divide="ignore", invalid="ignore"):
                t1 = (lo - oo) / dd; t2 = (hi - oo) / dd
            ta = np.minimum(t1, t2); tb = np.maximum(t1, t2)
            par = dd == 0
            inside = (oo >= lo) & (oo <= hi)
            ta = np.where(par, np.where(inside, -np.inf, np.inf), ta)
            tb = np.where(par, np.where(inside, np.inf, -np.inf), tb)
            tmin = np.maximum(tmin, ta); tmax = np.minimum(tmax, tb)
        hit = (tmax >= tmin) & (tmin > 1e-6)
        best = np.where(hit & (tmin < best), tmin, best)
    # poles (vertical cylinders)
    for cx, cy, r, h in scene.poles:
        ox, oy = o[0] - cx, o[1] - cy
        a = d[:, 0] ** 2 + d[:, 1] ** 2
        b = 2 * (ox * d[:, 0] + oy * d[:, 1])
        cc = ox * ox + oy * oy - r * r
        disc = b * b - 4 * a * cc
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a)
        z = o[2] + t * d[:, 2] - scene.ground_z
        hit = (disc > 0) & (a > 1e-12) & (t > 1e-6) & (z >= 0) & (z <= h)
        best = np.where(hit & (t < best), t, best)
    best = np.where(best > max_range, np.inf, best)
    return best


def simulate_scan(scene, R, t, model="VLP-16", noise=0.01, dropout=0.005, seed=0, rows=None, cols=None, tilt_deg=None,
                  max_range=None):
    """One organised frame (rows, cols, 4) float32 in the sensor frame; missing returns are NaN."""
    rng = np.random.default_rng(seed)
    dirs = ray_dirs(model, rows, cols, tilt_deg)
    rws, cls = dirs.shape[:2]
    if max_range is None:
        max_range = LIDARS[model][4] if tilt_deg is None else 12.0
    dw = dirs.reshape(-1, 3) @ np.asarray(R, np.float64).T
    rng_m = _cast(scene, np.asarray(t, np.float64), dw, max_range)
    rng_m = rng_m + rng.normal(0.0, noise, rng_m.shape)
    drop = rng.random(rng_m.shape) < dropout
    rng_m = np.where(drop, np.inf, rng_m)
    pts = dirs.reshape(-1, 3) * rng_m[:, None]
    out = np.empty((rws * cls, 4), np.float32)
    out[:, :3] = np.where(np.isfinite(rng_m)[:, None], pts, np.nan).astype(np.float32)
    out[:, 3] = rng.integers(1, 255, rws * cls).astype(np.float32)
    return out.reshape(rws, cls, 4)


def organised_to_sweep(frame):
    """Azimuth-major raw sweep (N, 4) as a spinning multi-beam driver emits it (all rings of column 0, then column 1, ...)."""
    rows, cols = frame.shape[:2]
    return np.ascontiguousarray(frame.transpose(1, 0, 2).reshape(rows * cols, 4))


def pose_matrix(yaw=0.0, pitch=0.0, roll=0.0, t=(0, 0, 0)):
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]]); Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz @ Ry @ Rx, np.asarray(t, np.float64)


def trajectory(n, seed=0, speed=1.0, z=0.0, yaw_amp=0.1):
    """n poses along the street (x axis): forward `speed` m per frame with a yaw sinusoid."""
    out = []
    for k in range(n):
        yaw = yaw_amp * np.sin(0.3 * k)
        out.append(pose_matrix(yaw, 0.01 * np.sin(0.7 * k), 0.01 * np.cos(0.5 * k), (speed * k - speed * n / 2, 0.3 * np.sin(0.2 * k), z)))
    return out


def sample_map(scene, spacing=0.4, seed=0, jitter=0.3, region=None):
    """Sample the scene surfaces (surf map) and its edges / poles (corner map) on a jittered lattice.

    Returns (corner (Nc,4), surf (Ns,4)) float32 in the world frame.  region = (xmin, xmax, ymin, ymax) crops.
    """
    rng = np.random.default_rng(seed)
    E = scene.extent * 1.5
    xmin, xmax, ymin, ymax = region if region is not None else (-E, E, -E, E)

    def lattice(u0, u1, v0, v1):
        nu = max(int((u1 - u0) / spacing), 1); nv = max(int((v1 - v0) / spacing), 1)
        u = u0 + (np.arange(nu) + 0.5) * (u1 - u0) / nu; v = v0 + (np.arange(nv) + 0.5) * (v1 - v0) / nv
        U, V = np.meshgrid(u, v, indexing="ij")
        U = U + rng.uniform(-jitter, jitter, U.shape) * spacing; V = V + rng.uniform(-jitter, jitter, V.shape) * spacing
        return U.ravel(), V.ravel()

    surf, corner = [], []
    gx, gy = lattice(xmin, xmax, ymin, ymax)
    keep = np.ones(gx.shape, bool)
    for cx, cy, yaw, bx, by, z0, z1 in scene.boxes:
        c, s = np.cos(yaw), np.sin(yaw)
        lx = c * (gx - cx) + s * (gy - cy); ly = -s * (gx - cx) + c * (gy - cy)
        keep &= ~((np.abs(lx) < bx) & (np.abs(ly) < by))
    surf.append(np.stack([gx[keep], gy[keep], np.full(keep.sum(), scene.ground_z)], -1))
    for cx, cy, yaw, bx, by, z0, z1 in scene.boxes:
        c, s = np.cos(yaw), np.sin(yaw)
        Rm = np.array([[c, -s], [s, c]])
        for axis, sign in ((0, 1), (0, -1), (1, 1), (1, -1)):   # four walls
            half = by if axis == 0 else bx
            u, v = lattice(-half, half, z0, z1)
            if axis == 0:
                loc = np.stack([np.full_like(u, sign * bx), u], -1)
            else:
                loc = np.stack([u, np.full_like(u, sign * by)], -1)
            w = loc @ Rm.T + np.array([cx, cy])
            surf.append(np.stack([w[:, 0], w[:, 1], v], -1))
        u, v = lattice(-bx, bx, -by, by)   # roof
        w = np.stack([u, v], -1) @ Rm.T + np.array([cx, cy])
        surf.append(np.stack([w[:, 0], w[:, 1], np.full(len(u), z1)], -1))
        for sx in (-1, 1):   # vertical edges
            for sy in (-1, 1):
                n = max(int((z1 - z0) / spacing), 1)
                zz = z0 + (np.arange(n) + 0.5) * (z1 - z0) / n + rng.uniform(-jitter, jitter, n) * spacing
                w = np.array([sx * bx, sy * by]) @ Rm.T + np.array([cx, cy])
                corner.append(np.stack([np.full(n, w[0]), np.full(n, w[1]), zz], -1))
        for axis in (0, 1):   # roof edges
            for sgn in (-1, 1):
                half = bx if axis == 0 else by
                n = max(int(2 * half / spacing), 1)
                uu = -half + (np.arange(n) + 0.5) * 2 * half / n + rng.uniform(-jitter, jitter, n) * spacing
                loc = np.stack([uu, np.full(n, sgn * by)], -1) if axis == 0 else np.stack([np.full(n, sgn * bx), uu], -1)
                w = loc @ Rm.T + np.array([cx, cy])
                corner.append(np.stack([w[:, 0], w[:, 1], np.full(n, z1)], -1))
    for cx, cy, r, h in scene.poles:
        n = max(int(h / spacing), 1)
        zz = scene.ground_z + (np.arange(n) + 0.5) * h / n + rng.uniform(-jitter, jitter, n) * spacing
        corner.append(np.stack([np.full(n, cx) + rng.normal(0, 0.02, n), np.full(n, cy) + rng.normal(0, 0.02, n), zz], -1))

    def pack(chunks):
        p = np.concatenate(chunks, 0) if chunks else np.zeros((0, 3))
        m = (p[:, 0] >= xmin) & (p[:, 0] <= xmax) & (p[:, 1] >= ymin) & (p[:, 1] <= ymax)
        p = p[m]
        out = np.zeros((len(p), 4), np.float32)
        out[:, :3] = p
        out[:, 3] = rng.uniform(0, 64, len(p)).astype(np.float32)
        return out

    return pack(corner), pack(surf)
