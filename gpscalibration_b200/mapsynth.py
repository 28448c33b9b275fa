"""Synthetic large maps and query stacks for the big-map configurations (BASELINE configs[2], [4]; SURVEY 8d cfg 3 / cfg 5).
Workload generators only (numpy): neither the oracle nor the CUDA path depends on them."""
import numpy as np


def synth_map_unfiltered(n_surf, n_corner, extent, seed=7):
    """Worst case (--unfiltered-map): uniformly random points on the planes / edges, NOT voxel-filtered, so wall cells hold
    ~11 and edge cells ~30 points (a real LM map never does: every cube is re-filtered at 0.2 / 0.4 m, LM:642-662)."""
    rng = np.random.default_rng(seed)
    ng = int(n_surf * 0.6)
    g = np.empty((ng, 4), np.float32)
    g[:, 0] = rng.uniform(-extent, extent, ng); g[:, 2] = rng.uniform(-extent / 2, extent / 2, ng)
    g[:, 1] = -1.8 + rng.normal(0, 0.01, ng); g[:, 3] = 0
    nw = n_surf - ng
    w = np.empty((nw, 4), np.float32)
    w[:, 0] = rng.uniform(-extent, extent, nw); w[:, 1] = rng.uniform(-1.8, 10, nw)
    w[:, 2] = (rng.integers(-6, 7, nw) * 40.0 + rng.normal(0, 0.01, nw)).astype(np.float32); w[:, 3] = 0
    c = np.empty((n_corner, 4), np.float32)
    c[:, 0] = rng.integers(-int(extent / 10), int(extent / 10) + 1, n_corner) * 10.0 + rng.normal(0, 0.01, n_corner)
    c[:, 2] = rng.integers(-6, 7, n_corner) * 40.0 + rng.normal(0, 0.01, n_corner)
    c[:, 1] = rng.uniform(-1.8, 10, n_corner); c[:, 3] = 0
    return c, np.concatenate([g, w]), extent


def synth_map(n_surf, n_corner, seed=7):
    """Planar 'city' as LM keeps it: every surface carries one point per 0.4 m voxel (surf) and every vertical edge one
    point per 0.2 m voxel (corner), i.e. what the per-cube voxel filter (LM:642-662) leaves.  Ground plane 2E x E,
    13 wall planes across it, edges on the walls; E follows from the requested point counts.  Returns (corner, surf, E)."""
    rng = np.random.default_rng(seed)
    n_g = int(n_surf * 0.8)
    nz = int(np.sqrt(n_g / 2)); nx = 2 * nz
    E = 0.4 * nz
    gx, gz = np.meshgrid(np.arange(nx, dtype=np.float32), np.arange(nz, dtype=np.float32), indexing="ij")
    g = np.empty((nx * nz, 4), np.float32)
    g[:, 0] = (gx.ravel() + rng.uniform(0.1, 0.9, nx * nz)) * 0.4 - E
    g[:, 2] = (gz.ravel() + rng.uniform(0.1, 0.9, nx * nz)) * 0.4 - E / 2
    g[:, 1] = -1.8 + rng.normal(0, 0.01, nx * nz); g[:, 3] = 0
    n_w = n_surf - nx * nz
    ny = max(1, n_w // (13 * nx))
    wx, wy, wk = np.meshgrid(np.arange(nx, dtype=np.float32), np.arange(ny, dtype=np.float32), np.arange(13, dtype=np.float32), indexing="ij")
    m = nx * ny * 13
    w = np.empty((m, 4), np.float32)
    w[:, 0] = (wx.ravel() + rng.uniform(0.1, 0.9, m)) * 0.4 - E
    w[:, 1] = (wy.ravel() + rng.uniform(0.1, 0.9, m)) * 0.4 - 1.8
    w[:, 2] = (wk.ravel() - 6) * (E / 13.0) + rng.normal(0, 0.01, m); w[:, 3] = 0
    H = ny * 0.4
    nyc = max(1, int(H / 0.2))
    nl = max(1, n_corner // (13 * nyc))  # edges per wall
    cx, cy, ck = np.meshgrid(np.arange(nl, dtype=np.float32), np.arange(nyc, dtype=np.float32), np.arange(13, dtype=np.float32), indexing="ij")
    m = nl * nyc * 13
    c = np.empty((m, 4), np.float32)
    c[:, 0] = (cx.ravel() + 0.5) * (2 * E / nl) - E + rng.normal(0, 0.01, m)
    c[:, 1] = (cy.ravel() + rng.uniform(0.1, 0.9, m)) * 0.2 - 1.8
    c[:, 2] = (ck.ravel() - 6) * (E / 13.0) + rng.normal(0, 0.01, m); c[:, 3] = 0
    return c, np.concatenate([g, w]), E



def voxel_order(c, leaf):
    """Stack clouds are voxel-grid outputs: ascending (k, j, i) cell order (PCL's linear cell id)."""
    ijk = np.floor(c[:, :3] / leaf).astype(np.int64)
    return c[np.lexsort((ijk[:, 0], ijk[:, 1], ijk[:, 2]))]


def synth_queries(corner_map, surf_map, n_queries, T_true, seed=11, ordered=True, take_all=False):
    """Query stacks = a random subset of the map moved by the inverse of the rigid transform T_true = {rx, ry, rz, tx, ty, tz}
    (pointAssociateTobeMapped, LM:264-282, in float64), so that Gauss-Newton started at zero converges towards T_true.
    Returns (corner_stack, surf_stack) in voxel-grid order (or random order)."""
    rng = np.random.default_rng(seed)
    if take_all:  # the caller has chosen the points
        qi_c, qi_s = np.arange(len(corner_map)), np.arange(len(surf_map))
    else:
        n_c = min(n_queries // 5, len(corner_map) // 2)
        qi_c = rng.choice(len(corner_map), n_c, replace=False)
        qi_s = rng.choice(len(surf_map), min(n_queries - n_c, len(surf_map)), replace=False)
    rx, ry, rz, tx, ty, tz = (float(v) for v in T_true)

    def inv(p):
        q = p.astype(np.float64)
        x1 = np.cos(ry) * (q[:, 0] - tx) - np.sin(ry) * (q[:, 2] - tz)
        y1 = q[:, 1] - ty
        z1 = np.sin(ry) * (q[:, 0] - tx) + np.cos(ry) * (q[:, 2] - tz)
        x2, y2, z2 = x1, np.cos(rx) * y1 + np.sin(rx) * z1, -np.sin(rx) * y1 + np.cos(rx) * z1
        out = p.copy()
        out[:, 0] = np.cos(rz) * x2 + np.sin(rz) * y2
        out[:, 1] = -np.sin(rz) * x2 + np.cos(rz) * y2
        out[:, 2] = z2
        return out

    cs, ss = inv(corner_map[qi_c]), inv(surf_map[qi_s])
    if ordered:
        cs, ss = voxel_order(cs, 0.2), voxel_order(ss, 0.4)
    return cs, ss


def synth_queries_local(corner_map, surf_map, n_corner, n_surf, T_true, center=(0.0, 0.0), radius=80.0, seed=13):
    """Query stacks shaped like ONE sweep's (LM:736-747): map points within `radius` of a sensor position in the ground
    plane (x, z), n_corner + n_surf of them, moved by the inverse of T_true, in voxel-grid order."""
    rng = np.random.default_rng(seed)

    def pick(m, n):
        near = np.flatnonzero((m[:, 0] - center[0]) ** 2 + (m[:, 2] - center[1]) ** 2 < radius * radius)
        return m[rng.choice(near, min(n, len(near)), replace=False)]

    sub_c, sub_s = pick(corner_map, n_corner), pick(surf_map, n_surf)
    cs, ss = synth_queries(sub_c, sub_s, 0, T_true, ordered=True, take_all=True)
    return cs, ss
