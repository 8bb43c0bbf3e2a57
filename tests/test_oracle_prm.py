"""Host-side checks of the oracle's PRM-build pieces: the grid k-NN against the O(n^2)
definition, and the batched sampleMultiple of the restatement against the reference's own
collision code (oracle/_ref) where it is built."""
import os
import tempfile

import numpy as np
import pytest

from ctopprm_learn_b200 import synth

CLR, CDC = synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK


@pytest.mark.parametrize("n,k", [(70, 13), (2500, 13), (4000, 5), (3000, 16)])
def test_knn_grid_equals_bruteforce(orc, n, k):
    rng = np.random.default_rng(n + k)
    pts = rng.random((n, 3)) * np.array([12.0, 9.0, 3.0])
    pts[10:30] = pts[10]                 # duplicates: (distance, index) tie-break
    if n == 4000:
        pts[:, 2] = 1.25                 # planar
    i0, d0 = orc.knn(pts, k, nthreads=4)
    i1, d1 = orc.knn_grid(pts, k, nthreads=4)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


def test_knn_grid_lattice_ties(orc):
    # regular lattice: many exactly equal distances
    g = np.stack(np.meshgrid(np.arange(12.0), np.arange(12.0), np.arange(12.0), indexing="ij"), -1).reshape(-1, 3) * 0.25
    i0, d0 = orc.knn(g, 13, nthreads=4)
    i1, d1 = orc.knn_grid(g, 13, nthreads=4)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


def _queries(c, e, q, n, seed):
    s, g = synth.random_queries(c, e, q, seed)
    cand = np.stack([synth.ellipsoid_candidates(s[i], g[i], 4 * n, seed * 100 + i) for i in range(q)])
    return s, g, cand


def test_sample_multiple_batch_matches_single(orc, c1_map):
    vox, c, e = c1_map
    om = orc.OracleMap(vox, c, e)
    q, n = 3, 400
    s, g, cand = _queries(c, e, q, n, 9)
    out = orc.sample_multiple_batch(om, s, g, cand, n, CLR, CDC, nthreads=3, full=True)
    assert out["short"] == 0
    for i in range(q):
        nodes = om.sample_reject(s[i], g[i], cand[i], CLR, n)[0]
        assert np.array_equal(out["nodes"][i], nodes)
        want = om.build_prm(nodes, CLR, CDC)
        nnz = int(out["nnz"][i])
        assert nnz == len(want["col"]) and out["n_free"][i] == want["directed_free"].sum()
        assert np.array_equal(out["row_ptr"][i], want["row_ptr"])
        assert np.array_equal(out["col"][i, :nnz], want["col"]) and np.array_equal(out["w"][i, :nnz], want["w"])
        assert out["lookups"][i] == want["lookups"]


def test_reference_sample_multiple_equals_restatement(orc, c1_map):
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built")
    vox, c, e = c1_map
    fn = tempfile.mktemp(suffix=".npy")
    synth.write_map(fn, vox, c, e)
    ref = orc.RefMap(fn)
    om = orc.OracleMap(vox, c, e)
    q, n = 4, 500
    s, g, cand = _queries(c, e, q, n, 5)
    a = orc.sample_multiple_batch(om, s, g, cand, n, CLR, CDC, nthreads=4, full=True)
    b = ref.sample_multiple_batch(s, g, cand, n, CLR, CDC, nthreads=4, full=True)
    for key in ("n_filled", "n_free", "nnz", "nodes", "row_ptr"):
        assert np.array_equal(a[key], b[key]), key
    for i in range(q):
        nnz = int(a["nnz"][i])
        assert np.array_equal(a["col"][i, :nnz], b["col"][i, :nnz]) and np.array_equal(a["w"][i, :nnz], b["w"][i, :nnz])
    # candidate stream too short is reported, not fatal
    short = ref.sample_multiple_batch(s[:1], g[:1], cand[:1, :40], n, CLR, CDC)
    assert short["short"] == 1 and short["n_filled"][0] < n
    ref.close()
    os.unlink(fn)


def test_division_shortcut_identity_is_exhaustively_exact(tmp_path):
    """k_edge_flat computes idx/npts as fma(fma(-q0, npts, idx), r, q0) with r = 1/npts for npts <= 65536;
    tests/csrc/div_identity.c checks every (idx, npts) pair against the true IEEE quotient."""
    import subprocess
    src = os.path.join(os.path.dirname(__file__), "csrc", "div_identity.c")
    exe = str(tmp_path / "div_identity")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-o", exe, src, "-lm"])
    out = subprocess.check_output([exe, "65536"]).decode().split()
    assert int(out[0]) == 0 and int(out[1]) == 65536 * 65535 // 2 - 0


def test_multiply_shift_division_identity(tmp_path):
    """the CSR kernels split flat edge indices with floor(x / d) = (x * mul) >> sh (ctp_prm.cuh, ctp_div_make);
    tests/csrc/udiv_identity.c checks the identity at the boundary cases of every d up to 300 000 and on
    2*10^7 random (x, d) pairs below 2^31"""
    import subprocess
    src = os.path.join(os.path.dirname(__file__), "csrc", "udiv_identity.c")
    exe = str(tmp_path / "udiv_identity")
    subprocess.check_call(["gcc", "-O2", "-o", exe, src])
    out = subprocess.check_output([exe, "300000"]).decode().split()
    assert int(out[0]) == 0 and int(out[1]) > 2 * 10 ** 7


def test_oracle_sssp_against_scipy(orc, c1_map):
    """the oracle's Dijkstra (restating include/dijkstra.hpp:83-176) against scipy's on a PRM graph"""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra
    vox, c, e = c1_map
    om = orc.OracleMap(vox, c, e)
    s, g, cand = _queries(c, e, 1, 1500, 3)
    nodes = om.sample_reject(s[0], g[0], cand[0], CLR, 1500)[0]
    gph = om.build_prm(nodes, CLR, CDC, nthreads=4)
    dist, pred = orc.sssp(gph["row_ptr"], gph["col"], gph["w"], [0])
    ref = dijkstra(csr_matrix((gph["w"], gph["col"], gph["row_ptr"]), shape=(1500, 1500)), directed=True, indices=0)
    assert np.array_equal(np.isfinite(dist), np.isfinite(ref))
    fin = np.isfinite(ref)
    assert np.allclose(dist[fin], ref[fin], rtol=1e-12, atol=0)
    # predecessor chains reproduce the distances
    for v in np.flatnonzero(fin)[::37]:
        total, u = 0.0, v
        while pred[u] >= 0:
            p = pred[u]
            row = gph["col"][gph["row_ptr"][u]:gph["row_ptr"][u + 1]]
            total += gph["w"][gph["row_ptr"][u] + int(np.flatnonzero(row == p)[0])]
            u = p
        assert dist[u] == 0.0 and abs(total - dist[v]) <= 1e-9 * max(1.0, dist[v])
    # multi-source flood: label = nearest source
    srcs = [0, 1, 700]
    dm, pm, lab = orc.sssp(gph["row_ptr"], gph["col"], gph["w"], srcs, want_label=True)
    each = np.stack([orc.sssp(gph["row_ptr"], gph["col"], gph["w"], [x])[0] for x in srcs])
    assert np.allclose(dm[np.isfinite(dm)], each.min(0)[np.isfinite(dm)], rtol=1e-12)
    clear = np.isfinite(dm) & (np.sort(each, 0)[1] - np.sort(each, 0)[0] > 1e-9)
    assert np.array_equal(lab[clear], each.argmin(0)[clear])


def test_edt_oracle_matches_scipy_and_definition(orc):
    """orc_edt_signed (checker of the device distance transform) against scipy.ndimage.distance_transform_edt and,
    on a tiny grid, against the O(n^6) definition"""
    ndi = pytest.importorskip("scipy.ndimage")
    for n, n_box, fill, seed in ((6, 1, 0.05, 1), (24, 5, 0.001, 2), (40, 12, None, 3)):
        occ = synth.blocks_occupancy(n, n_box, seed, fill)
        vox, d2 = orc.edt_signed(occ, 0.1)
        out = ndi.distance_transform_edt(occ == 0)
        ins = ndi.distance_transform_edt(occ != 0)
        ref = np.rint(out ** 2).astype(np.int64) - np.rint(ins ** 2).astype(np.int64)
        assert np.array_equal(d2, ref)
        want = np.where(d2 >= 0, 0.1 * np.sqrt(np.abs(d2).astype(np.float64)), -(0.1 * np.sqrt(np.abs(d2).astype(np.float64))))
        assert np.array_equal(vox, want.astype(np.float32))
        if n == 6:
            idx = np.argwhere(np.ones_like(occ)).astype(np.int64)
            dd = ((idx[:, None, :] - idx[None, :, :]) ** 2).sum(-1)
            o = occ.reshape(-1) != 0
            brute = np.where(o, -np.where(~o[None, :], dd, 1 << 40).min(1), np.where(o[None, :], dd, 1 << 40).min(1))
            assert np.array_equal(d2.reshape(-1), brute)
    # no voxel of the other kind: the documented sentinel 3 n^2
    _, d2 = orc.edt_signed(np.zeros((5, 5, 5), np.uint8), 1.0)
    assert (d2 == 75).all()
    _, d2 = orc.edt_signed(np.ones((5, 5, 5), np.uint8), 1.0)
    assert (d2 == -75).all()
