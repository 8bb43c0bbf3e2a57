"""Synthetic inputs for the CTopPRM hot path (numpy only; no reference code involved).

Maps follow the on-disk format written by the reference's ``map.py:124-132`` and read by
``src/esdf_map.cpp:15-25``: four back-to-back NPY records (voxels ``(N,N,N)`` f4, centroid
``(3,)`` f8, extents ``(3,)`` f8, voxel_resolution ``()`` i8).  Voxel ``i`` along an axis sits
at ``centroid - E/2 + i*E/N`` -- the inverse of the world->index transform in
``src/esdf_map.cpp:136-139`` (note ``N/2``, not ``(N-1)/2``).

Shapes and seeds are the ones fixed in BASELINE.md section 4 / SURVEY.md section 8(d).
"""
from __future__ import annotations

import numpy as np

# BASELINE.md section 4 (builder-authored defaults; the reference ships no YAML)
MIN_CLEARANCE = 0.2
COLLISION_DISTANCE_CHECK = 0.05
ELLIPSE_RATIO = 1.2
NUM_NN = 14  # include/topological_prm_clustering.hpp:291 (13 neighbours + self)

CONFIGS = {
    # id: (kind, extent_m, voxels_per_axis, n_obstacles, seed, nodes)
    "C1": ("columns", 12.8, 128, 40, 1, 1000),
    "C2": ("columns", 12.8, 256, 60, 2, 10000),
    # 2000 trunks (+ 1000 canopies): with the ~4000 of the survey's sketch the trunks, grown by min_clearance, cover the
    # ground past the percolation threshold (largest free component: 5 % of the free area, nothing spans the map), so no
    # query longer than a few metres has a path and the distinct-path stage of configs[2] has nothing to find; at 2000
    # the free space is one component (98 %) that spans the forest.  Recorded in BASELINE.md section 4.
    "C3": ("forest", 51.2, 1024, 2000, 3, 100000),
}


def centroid_for(extent: float) -> np.ndarray:
    # deliberately non-dyadic: keeps lookups away from the integer-index 0/0 quirk
    return np.array([0.013, -0.007, extent / 2 + 0.011], dtype=np.float64)


def axis_coords(centroid: np.ndarray, extent: float, n: int, axis: int) -> np.ndarray:
    return centroid[axis] - extent / 2 + np.arange(n, dtype=np.float64) * (extent / n)


def random_columns_esdf(n: int, extent: float, n_col: int, seed: int):
    """Vertical cylinders, analytic signed distance (negative inside), constant along z."""
    rng = np.random.default_rng(seed)
    c = centroid_for(extent)
    cx = c[0] + (rng.random(n_col) - 0.5) * extent
    cy = c[1] + (rng.random(n_col) - 0.5) * extent
    r = 0.15 + rng.random(n_col) * 0.45
    x = axis_coords(c, extent, n, 0)
    y = axis_coords(c, extent, n, 1)
    vox = columns_field(n, c, extent, cx, cy, r)
    return vox, c, np.array([extent] * 3, dtype=np.float64)


def column_params(n_col: int, extent: float, seed: int):
    """centres and radii of the random columns of seed `seed` (what random_columns_esdf draws)"""
    rng = np.random.default_rng(seed)
    c = centroid_for(extent)
    cx = c[0] + (rng.random(n_col) - 0.5) * extent
    cy = c[1] + (rng.random(n_col) - 0.5) * extent
    r = 0.15 + rng.random(n_col) * 0.45
    return c, cx, cy, r


def columns_field(n: int, c, extent: float, cx, cy, r):
    """signed distance to a set of vertical cylinders, min_i(sqrt(dx^2 + dy^2) - r_i), FP64 with one rounding
    per operation (the order ctp_map_create_columns evaluates it in, so both give the same floats)"""
    x = axis_coords(c, extent, n, 0)
    y = axis_coords(c, extent, n, 1)
    d2 = np.full((n, n), np.inf)
    for i in range(len(cx)):
        dx = x[:, None] - cx[i]
        dy = y[None, :] - cy[i]
        d = np.sqrt(dx * dx + dy * dy) - r[i]
        np.minimum(d2, d, out=d2)
    return np.ascontiguousarray(np.broadcast_to(d2[:, :, None], (n, n, n)), dtype=np.float32)


def forest_esdf(n: int, extent: float, n_trunk: int, seed: int, n_canopy: int | None = None,
                canopy_reach: float = 1.5):
    """Trunks (vertical cylinders r~U[0.1,0.5]) plus spherical canopies.  Canopy distance is
    only stamped within ``canopy_reach`` metres of each sphere (enough for any clearance
    threshold used here); the field stays a valid lower-bounded clearance everywhere."""
    rng = np.random.default_rng(seed)
    c = centroid_for(extent)
    if n_canopy is None:
        n_canopy = n_trunk // 2
    cx = c[0] + (rng.random(n_trunk) - 0.5) * extent
    cy = c[1] + (rng.random(n_trunk) - 0.5) * extent
    r = 0.1 + rng.random(n_trunk) * 0.4
    x = axis_coords(c, extent, n, 0)
    y = axis_coords(c, extent, n, 1)
    z = axis_coords(c, extent, n, 2)
    d2 = np.full((n, n), np.inf, dtype=np.float32)
    h = extent / n
    for i in range(n_trunk):  # local stamping keeps this O(n_trunk * window)
        reach = r[i] + 4.0
        i0, i1 = np.searchsorted(x, [cx[i] - reach, cx[i] + reach])
        j0, j1 = np.searchsorted(y, [cy[i] - reach, cy[i] + reach])
        if i1 <= i0 or j1 <= j0:
            continue
        d = (np.hypot(x[i0:i1, None] - cx[i], y[None, j0:j1] - cy[i]) - r[i]).astype(np.float32)
        np.minimum(d2[i0:i1, j0:j1], d, out=d2[i0:i1, j0:j1])
    d2 = np.minimum(d2, np.float32(4.0))
    vox = np.empty((n, n, n), dtype=np.float32)
    vox[:] = d2[:, :, None]
    k = rng.integers(0, n_trunk, n_canopy)
    sz = c[2] + (rng.random(n_canopy) - 0.5) * extent * 0.6
    sr = 0.5 + rng.random(n_canopy) * 1.0
    for i in range(n_canopy):
        reach = sr[i] + canopy_reach
        i0, i1 = np.searchsorted(x, [cx[k[i]] - reach, cx[k[i]] + reach])
        j0, j1 = np.searchsorted(y, [cy[k[i]] - reach, cy[k[i]] + reach])
        k0, k1 = np.searchsorted(z, [sz[i] - reach, sz[i] + reach])
        if i1 <= i0 or j1 <= j0 or k1 <= k0:
            continue
        d = np.sqrt((x[i0:i1, None, None] - cx[k[i]]) ** 2 + (y[None, j0:j1, None] - cy[k[i]]) ** 2
                    + (z[None, None, k0:k1] - sz[i]) ** 2) - sr[i]
        sub = vox[i0:i1, j0:j1, k0:k1]
        np.minimum(sub, d.astype(np.float32), out=sub)
    del h
    return vox, c, np.array([extent] * 3, dtype=np.float64)


def make_config_map(cfg: str, n_override: int | None = None):
    kind, extent, n, n_obs, seed, _ = CONFIGS[cfg]
    if n_override:
        n = n_override
    if kind == "columns":
        return random_columns_esdf(n, extent, n_obs, seed)
    return forest_esdf(n, extent, n_obs, seed)


def write_map(path, vox, centroid, extents, resolution=None):
    """map.py:124-132 -- four np.save calls into one file."""
    with open(path, "wb") as f:
        np.save(f, np.ascontiguousarray(vox))
        np.save(f, np.asarray(centroid, dtype=np.float64))
        np.save(f, np.asarray(extents, dtype=np.float64))
        np.save(f, np.int64(vox.shape[0] if resolution is None else resolution))


def read_map(path):
    with open(path, "rb") as f:
        vox = np.load(f)
        centroid = np.load(f)
        extents = np.load(f)
        res = np.load(f)
    return vox, centroid, extents, res


def default_query(centroid, extents, cfg=None):
    """start/goal on opposite x faces, inset 1 m, mid height.  With cfg (a CONFIGS key) each end slides
    along its face (y, in 0.25 m steps, alternating sides) until it lies in free space of that map."""
    c = np.asarray(centroid, dtype=np.float64)
    e = np.asarray(extents, dtype=np.float64)
    s = c.copy()
    g = c.copy()
    s[0] = c[0] - e[0] / 2 + 1.0
    g[0] = c[0] + e[0] / 2 - 1.0
    s[1] += 0.37
    g[1] -= 0.41
    if cfg is not None:
        for p in (s, g):
            y0 = p[1]
            for step in range(int(e[1] / 0.25)):
                p[1] = y0 + 0.25 * ((step + 1) // 2) * (1 if step % 2 else -1)
                if abs(p[1] - c[1]) < e[1] / 2 - 1.0 and analytic_clearance(cfg, p[None])[0] >= QUERY_CLEARANCE:
                    break
    return s, g


def config_obstacles(cfg: str):
    """the analytic obstacles of a configured map: (centroid, trunks (cx, cy, r), canopies (x, y, z, r) or None) --
    the same draws, in the same order, as random_columns_esdf / forest_esdf make"""
    kind, extent, n, n_obs, seed, _ = CONFIGS[cfg]
    rng = np.random.default_rng(seed)
    c = centroid_for(extent)
    cx = c[0] + (rng.random(n_obs) - 0.5) * extent
    cy = c[1] + (rng.random(n_obs) - 0.5) * extent
    if kind == "columns":
        r = 0.15 + rng.random(n_obs) * 0.45
        return c, (cx, cy, r), None
    r = 0.1 + rng.random(n_obs) * 0.4
    n_canopy = n_obs // 2
    k = rng.integers(0, n_obs, n_canopy)
    sz = c[2] + (rng.random(n_canopy) - 0.5) * extent * 0.6
    sr = 0.5 + rng.random(n_canopy) * 1.0
    return c, (cx, cy, r), (cx[k], cy[k], sz, sr)


def analytic_clearance(cfg, pts):
    """distance from pts (m x 3) to the nearest analytic obstacle of a configured map (not the gridded field: used
    only to draw start / goal positions in free space).  cfg: a CONFIGS key, or the
    obstacles themselves as ((cx, cy, r), canopies or None)."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
    if isinstance(cfg, str):
        _, (cx, cy, r), canopy = config_obstacles(cfg)
    else:
        (cx, cy, r), canopy = cfg
    out = np.full(len(pts), np.inf)
    for lo in range(0, len(pts), 4096):
        p = pts[lo:lo + 4096]
        d = np.hypot(p[:, None, 0] - cx[None], p[:, None, 1] - cy[None]) - r[None]
        best = d.min(axis=1)
        if canopy is not None:
            sx, sy, sz, sr = canopy
            ds = np.sqrt((p[:, None, 0] - sx[None]) ** 2 + (p[:, None, 1] - sy[None]) ** 2 + (p[:, None, 2] - sz[None]) ** 2) - sr[None]
            best = np.minimum(best, ds.min(axis=1))
        out[lo:lo + 4096] = best
    return out


# start / goal of the C4 query set keep this much distance to the analytic obstacles (min_clearance + 2 voxels of
# slack for the gridded field): "start-goal pairs ~U in free space" 
QUERY_CLEARANCE = MIN_CLEARANCE + 0.1


def random_queries(centroid, extents, q: int, seed: int, min_sep_frac: float = 1 / 3, cfg=None):
    """C4: start-goal pairs ~U inside the box (1 m inset) with ||g-s|| >= E*min_sep_frac; with cfg (a CONFIGS key)
    both ends are also required to lie in free space of that map (analytic obstacles, QUERY_CLEARANCE)."""
    rng = np.random.default_rng(seed)
    c = np.asarray(centroid, dtype=np.float64)
    e = np.asarray(extents, dtype=np.float64)
    out_s = np.empty((q, 3))
    out_g = np.empty((q, 3))
    k = 0
    while k < q:
        # pairs are drawn in blocks so that the free-space test is vectorised; the draw order does not depend on q
        blk = 256
        s = c + (rng.random((blk, 3)) - 0.5) * (e - 2.0)
        g = c + (rng.random((blk, 3)) - 0.5) * (e - 2.0)
        ok = np.linalg.norm(g - s, axis=1) >= e.max() * min_sep_frac
        if cfg is not None:
            ok &= (analytic_clearance(cfg, s) >= QUERY_CLEARANCE) & (analytic_clearance(cfg, g) >= QUERY_CLEARANCE)
        for i in np.nonzero(ok)[0]:
            if k < q:
                out_s[k], out_g[k] = s[i], g[i]
                k += 1
    return out_s, out_g


def ellipsoid_frame(start, goal):
    """Rotation with first column along goal-start; same completion rule as the oracle's
    restatement of src/common.cpp:256-312."""
    a1 = (goal - start) / np.linalg.norm(goal - start)
    k = int(np.argmin(np.abs(a1)))
    ek = np.zeros(3)
    ek[k] = 1.0
    a2 = np.cross(a1, ek)
    a2 /= np.linalg.norm(a2)
    a3 = np.cross(a1, a2)
    return np.stack([a1, a2, a3], axis=1)


def ellipsoid_candidates(start, goal, m: int, seed: int, ratio: float = ELLIPSE_RATIO,
                         planar: bool = False):
    """Candidate stream (pre-rejection) from the reference's spheroid distribution:
    1 uniform + 3 normal draws per point (src/common.cpp:295-303)."""
    start = np.asarray(start, dtype=np.float64)
    goal = np.asarray(goal, dtype=np.float64)
    rng = np.random.default_rng(seed)
    u = rng.random(m)
    g = rng.standard_normal((m, 3))
    two_focal = np.linalg.norm(goal - start)
    two_major = two_focal * ratio
    axis = np.sqrt(two_major ** 2 - two_focal ** 2)
    ball = g * (np.cbrt(u) / np.linalg.norm(g, axis=1))[:, None]
    scaled = ball * np.array([two_major / 2, axis / 2, axis / 2])
    pts = scaled @ ellipsoid_frame(start, goal).T + (start + goal) / 2
    if planar:
        pts[:, 2] = start[2]
    return np.ascontiguousarray(pts)


def random_edges(centroid, extents, m: int, seed: int, lo: float = 0.2, hi: float = 1.5):
    """Microbench edge list: m directed segments, uniform start, isotropic direction,
    length ~U[lo,hi]."""
    rng = np.random.default_rng(seed)
    c = np.asarray(centroid, dtype=np.float64)
    e = np.asarray(extents, dtype=np.float64)
    a = c + (rng.random((m, 3)) - 0.5) * (e - 0.2)
    d = rng.standard_normal((m, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    b = a + d * (lo + rng.random(m) * (hi - lo))[:, None]
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def blocks_occupancy(n, n_box, seed, fill=None):
    """seeded voxel occupancy grid: n_box axis-aligned boxes (and, with fill, random single voxels of that
    density) -- input of the device distance transform (ctp_map_create_occupancy)"""
    rng = np.random.default_rng(seed)
    occ = np.zeros((n, n, n), dtype=np.uint8)
    for _ in range(n_box):
        lo = rng.integers(0, n, 3)
        sz = rng.integers(1, max(2, n // 6), 3)
        hi = np.minimum(lo + sz, n)
        occ[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = 1
    if fill:
        occ |= (rng.random((n, n, n)) < fill).astype(np.uint8)
    return occ
