"""The float32 whole-edge filter of the roadmap-edge path (ctp_device.cuh, k_edge_mid) must never change a result: every
case here runs with the filter on and off and compares both with the oracle -- directed verdicts and the reference-order
lookup count -- bit for bit.  The inputs are built to sit where a float32 evaluation could go wrong: clearances equal to
the threshold or one ulp off it, samples exactly on voxel lattice planes (the reference's 0/0 -> NaN quirk), samples in
cell 0 and outside the grid, constant grids whose FP64 lerp wobbles around the threshold, grids with infinities."""
import numpy as np
import pytest

from ctopprm_learn_b200 import capi, synth

pytestmark = pytest.mark.gpu

CLR, CDC = synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK
LAYOUTS = (capi.LAYOUT_PLAIN, capi.LAYOUT_CELL8, capi.LAYOUT_TEX, capi.LAYOUT_BRICK)
ALL = capi.LAYOUT_PLAIN | capi.LAYOUT_CELL8 | capi.LAYOUT_TEX | capi.LAYOUT_BRICK


def _check(m, om, a, b, clr, cdc, layouts=LAYOUTS, mode=capi.EDGE_NODES):
    """the segments a -> b and b -> a as the k = 1 neighbour table of the 2m nodes (a, b): the roadmap-edge path
    (ctp_edge_check_knn_dev: whole-edge filter + exact kernel on what it leaves) with the filter on and off"""
    import torch
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 3)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1, 3)
    cnt = len(a)
    f_ab, _, _, l_ab, t_ab = om.edge_nodes(a, b, clr, cdc, nthreads=8)
    f_ba, _, _, l_ba, t_ba = om.edge_nodes(b, a, clr, cdc, nthreads=8)
    want = np.concatenate([f_ab, f_ba])
    dev = torch.device("cuda", m.device)
    nodes = torch.from_numpy(np.concatenate([a, b])).to(dev)
    nbr = torch.from_numpy(np.concatenate([np.arange(cnt, 2 * cnt), np.arange(0, cnt)]).astype(np.int32)).to(dev)
    for layout in layouts:
        m.set_layout(layout)
        for filt in (2, 0):  # 2 = always (1 would drop the filter where its calibration says it does not pay)
            m.set_tunable(capi.TUNE_EDGE_FILTER, filt)
            m.read_lookups()
            free = torch.full((2 * cnt,), 7, dtype=torch.uint8, device=dev)
            m.edge_check_knn_dev(nodes, 1, 2 * cnt, 1, nbr, clr, cdc, free)
            torch.cuda.synchronize(dev)
            got = free.cpu().numpy()
            assert set(np.unique(got)) <= {0, 1}, (layout, filt)
            assert np.array_equal(got.astype(bool), want), (layout, filt, int((got.astype(bool) != want).sum()))
            assert m.read_lookups() == t_ab + t_ba, (layout, filt)
    m.set_tunable(capi.TUNE_EDGE_FILTER, 1)
    return f_ab


def test_filter_equals_exact_on_random_segments(orc, c1_map):
    vox, c, e = c1_map
    m = capi.DeviceMap(vox, c, e, layouts=ALL)
    om = orc.OracleMap(vox, c, e)
    try:
        a, b = synth.random_edges(c, e, 300000, 11, lo=0.0, hi=3.0)
        f = _check(m, om, a, b, CLR, CDC)
        assert 0.05 < f.mean() < 0.95
        # segments that start / end outside the box, and ones longer than the filter accepts (> 256 voxels per axis)
        rng = np.random.default_rng(5)
        a2 = c + (rng.random((50000, 3)) - 0.5) * e * 1.3
        b2 = c + (rng.random((50000, 3)) - 0.5) * e * 1.3
        _check(m, om, a2, b2, CLR, CDC, layouts=(capi.LAYOUT_TEX, capi.LAYOUT_PLAIN))
        _check(m, om, a2, a2 + (b2 - a2) * 40.0, 0.05, 0.6, layouts=(capi.LAYOUT_TEX,))
    finally:
        m.close()


def test_filter_thresholds_taken_from_the_samples_themselves(orc, c1_map):
    """min_clearance set to the exact FP64 clearance of a sample (free at equality) and to its two neighbours"""
    vox, c, e = c1_map
    m = capi.DeviceMap(vox, c, e, layouts=capi.LAYOUT_TEX | capi.LAYOUT_PLAIN)
    om = orc.OracleMap(vox, c, e)
    try:
        a, b = synth.random_edges(c, e, 60000, 12, lo=0.1, hi=1.5)
        d = np.linalg.norm(b - a, axis=1)
        npts = np.ceil(d / CDC + 1.0)
        rng = np.random.default_rng(3)
        for _ in range(6):
            i = int(rng.integers(len(a)))
            idx = int(rng.integers(1, int(npts[i])))
            p = a[i] + (b[i] - a[i]) * (idx / npts[i])
            cl = float(om.clearance(p[None])[0])
            if not np.isfinite(cl):
                continue
            for thr in (cl, np.nextafter(cl, np.inf), np.nextafter(cl, -np.inf), float(np.float32(cl))):
                _check(m, om, a, b, thr, CDC, layouts=(capi.LAYOUT_TEX,))
    finally:
        m.close()


def test_filter_samples_on_lattice_planes_and_in_cell_zero(orc):
    """a map whose index transform is exact in binary (N = 32, extent 2, centroid 0: q = 16 p + 16): segments with
    coordinates on voxel planes hit the reference's 0/0 quirk at every sample, segments in cell 0 need the position test"""
    n = 32
    rng = np.random.default_rng(8)
    vox = (0.3 + 0.4 * rng.random((n, n, n))).astype(np.float32)
    c, e = np.zeros(3), np.full(3, 2.0)
    m = capi.DeviceMap(vox, c, e, layouts=ALL)
    om = orc.OracleMap(vox, c, e)
    try:
        k = 4000
        a = (rng.integers(1, n - 2, (k, 3)) - n / 2) / 16.0 + rng.random((k, 3)) / 16.0
        b = a + (rng.random((k, 3)) - 0.5) * 0.8
        # one or two coordinates exactly on a lattice plane, constant along the segment
        for ax in range(3):
            sel = slice(ax * 600, ax * 600 + 600)
            a[sel, ax] = np.round(a[sel, ax] * 16.0) / 16.0
            b[sel, ax] = a[sel, ax]
        a[1800:2100, :2] = np.round(a[1800:2100, :2] * 16.0) / 16.0
        b[1800:2100, :2] = a[1800:2100, :2]
        # segments along an axis between lattice points (every 16th of the samples lands on a lattice point)
        a[2100:2400] = np.round(a[2100:2400] * 16.0) / 16.0
        b[2100:2400] = a[2100:2400]
        b[2100:2400, 0] += 0.5
        # cell 0 and the last cell on each axis, just inside the box
        a[2400:2700, 0] = -1.0 + rng.random(300) / 16.0
        a[2700:3000, 1] = 1.0 - rng.random(300) / 16.0
        a[3000:3100, 2] = -1.0
        a[3100:3200, 2] = 1.0
        for cdc in (0.05, 1.0 / 32.0, 0.011):
            f = _check(m, om, a, b, 0.45, cdc)
            assert 0.02 < f.mean() < 0.98
    finally:
        m.close()


def test_filter_on_degenerate_grids(orc):
    """constant grid at the threshold (the FP64 lerp of eight equal corners is not always that constant), a grid with
    infinities / NaNs (the filter must switch itself off), a grid of huge values (band wider than the cut-off)"""
    n = 24
    c, e = np.array([0.3, -0.2, 0.1]), np.full(3, 3.0)
    rng = np.random.default_rng(4)
    a = c + (rng.random((20000, 3)) - 0.5) * e * 0.95
    b = a + (rng.random((20000, 3)) - 0.5) * 1.2
    thr = float(np.float32(0.2))
    grids = []
    grids.append((np.full((n, n, n), np.float32(0.2)), (thr, 0.2, np.nextafter(thr, np.inf))))
    g = (0.1 + 0.3 * rng.random((n, n, n))).astype(np.float32)
    g[5, 5, 5] = np.inf
    g[10, 3, 7] = -np.inf
    g[12:14, 12:14, 12:14] = np.nan
    grids.append((g, (0.2,)))
    grids.append(((1e30 * (rng.random((n, n, n)) - 0.3)).astype(np.float32), (0.2, 1e29)))
    grids.append(((1e-3 * rng.random((n, n, n))).astype(np.float32), (5e-4,)))
    for vox, thrs in grids:
        m = capi.DeviceMap(vox, c, e, layouts=capi.LAYOUT_TEX | capi.LAYOUT_PLAIN)
        om = orc.OracleMap(vox, c, e)
        try:
            for t in thrs:
                _check(m, om, a, b, t, 0.05, layouts=(capi.LAYOUT_TEX, capi.LAYOUT_PLAIN))
        finally:
            m.close()


def test_filter_far_from_the_origin(orc):
    """centroids so far out that the FP64 roundings of the reference's own index transform grow (eta): the filter widens
    its integer band or switches itself off, results stay the oracle's"""
    nvox = 40
    vox, c, e = synth.random_columns_esdf(nvox, 4.0, 6, 9)
    rng = np.random.default_rng(6)
    for shift in (250.0, 1.0e5, 3.0e7, -4.0e9):
        cc = c + shift
        m = capi.DeviceMap(vox, cc, e, layouts=capi.LAYOUT_TEX)
        om = orc.OracleMap(vox, cc, e)
        try:
            a = cc + (rng.random((30000, 3)) - 0.5) * e
            b = a + (rng.random((30000, 3)) - 0.5) * 1.5
            f = _check(m, om, a, b, 0.1, 0.05, layouts=(capi.LAYOUT_TEX,))
            assert 0.0 < f.mean() < 1.0
        finally:
            m.close()


def test_filter_in_the_prm_build(orc, c1_map):
    """the k-NN edges of a roadmap through ctp_build_prm with the filter on and off: same directed verdicts, same CSR"""
    vox, c, e = c1_map
    m = capi.DeviceMap(vox, c, e, layouts=capi.LAYOUT_TEX | capi.LAYOUT_PLAIN)
    om = orc.OracleMap(vox, c, e)
    try:
        s, g = synth.default_query(c, e)
        cand = synth.ellipsoid_candidates(s, g, 30000, 21)
        nodes = om.sample_reject(s, g, cand, CLR, 5000)[0]
        want = om.build_prm(nodes, CLR, CDC, nthreads=8)
        for filt in (2, 0):  # 2 = always (1 would drop the filter where its calibration says it does not pay)
            m.set_tunable(capi.TUNE_EDGE_FILTER, filt)
            out = m.build_prm(nodes[None], CLR, CDC, k=13)
            nnz = int(out["nnz_off"][1])
            assert np.array_equal(out["row_ptr"][0], want["row_ptr"])
            assert np.array_equal(out["col"][:nnz], want["col"])
            assert np.array_equal(out["w"][:nnz], want["w"])
            assert np.array_equal(out["directed_free"][0].astype(bool), want["directed_free"].astype(bool))
    finally:
        m.close()
