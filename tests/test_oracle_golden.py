"""The CPU restatement (oracle/ctp_oracle.c) against golden vectors produced by the
reference's own src/esdf_map.cpp + src/base_map.cpp (tools/make_golden.py -> oracle/_ref).
Bit-exact: FP64 values compared with array_equal, NaN positions identical."""
import numpy as np
import pytest


def _split(g, key_pts, key_off):
    off = g[key_off]
    return [g[key_pts][off[i]:off[i + 1]] for i in range(len(off) - 1)]


def test_bounds(golden, orc):
    om = orc.OracleMap(golden["vox"], golden["centroid"], golden["extents"])
    assert np.array_equal(om.min_pos(), golden["min_pos"])
    assert np.array_equal(om.max_pos(), golden["max_pos"])


def test_clearance_bit_exact(golden, orc):
    om = orc.OracleMap(golden["vox"], golden["centroid"], golden["extents"])
    got = om.clearance(golden["pts"])
    assert np.array_equal(np.isnan(got), np.isnan(golden["clearance"]))
    assert np.array_equal(got, golden["clearance"], equal_nan=True)
    # the integer-index quirk (0/0 -> NaN, src/esdf_map.cpp:108-110) is present in the vectors
    assert np.isnan(golden["clearance"][4000:4064]).sum() >= 48  # a few miss an exact integer by 1 ulp
    assert np.array_equal(om.in_collision(golden["pts"], float(golden["clr"])), golden["in_collision"])


def test_gradient_bit_exact(golden, orc):
    om = orc.OracleMap(golden["vox"], golden["centroid"], golden["extents"])
    g, c = om.gradient(golden["pts"][:1500])
    assert np.array_equal(g, golden["grad"], equal_nan=True)
    assert np.array_equal(c, golden["centre"], equal_nan=True)


def test_edges_bit_exact(golden, orc):
    om = orc.OracleMap(golden["vox"], golden["centroid"], golden["extents"])
    clr, cdc = float(golden["clr"]), float(golden["cdc"])
    free, hit, hit_idx, lookups, _ = om.edge_nodes(golden["a"], golden["b"], clr, cdc)
    assert np.array_equal(free, golden["free_nodes"])
    assert np.array_equal(hit, golden["hit"], equal_nan=True)
    assert (hit_idx[free] == -1).all() and (hit_idx[~free] >= 1).all()
    assert np.array_equal(lookups[~free], hit_idx[~free])
    fg, _ = om.edge_guards(golden["a"], golden["b"], clr, cdc)
    assert np.array_equal(fg, golden["free_guards"])
    f2 = om.edge_nodes(golden["a"], golden["b"], 0.05, 0.11)[0]
    assert np.array_equal(f2, golden["free_nodes_alt"])
    # nodes/guards differ exactly where the d < cdc shortcut or direction matters
    assert free[:8].all()


def test_sample_path_bit_exact(golden, orc):
    polys = _split(golden, "poly_pts", "poly_off")
    pos = 0
    for p, L, k in zip(polys, golden["poly_len"], golden["poly_num"]):
        assert orc.path_length(p) == L
        out, rc = orc.sample_path(p, L, k)
        assert rc == 0
        assert np.array_equal(out, golden["sampled"][pos:pos + int(k)])
        pos += int(k)


def test_deformable(golden, orc):
    om = orc.OracleMap(golden["vox"], golden["centroid"], golden["extents"])
    cdc = float(golden["cdc"])
    polys = _split(golden, "poly_pts", "poly_off")
    lens = golden["poly_len"]
    for i, j, want in zip(golden["pair_i"], golden["pair_j"], golden["deformable"]):
        k = float(np.ceil(max(lens[i], lens[j]) / cdc) + 1)
        assert om.deformable(polys[i], lens[i], polys[j], lens[j], k, 0.02, cdc) == want
    near = golden["near_pts"].reshape(3, -1, 3)
    nl = golden["near_len"]
    idx = 0
    for i in range(3):
        for j in range(i + 1, 3):
            k = float(np.ceil(max(nl[i], nl[j]) / cdc) + 1)
            assert om.deformable(near[i], nl[i], near[j], nl[j], k, -10.0, cdc) == golden["near_deformable"][idx]
            idx += 1
    assert golden["near_deformable"].any()
    q = golden["quad"]
    assert om.deformable_between(q[0], q[1], q[2], q[3], -10.0, cdc) == golden["deformable_between"][0]
    assert om.deformable_between(q[0], q[1], q[2], q[3], float(golden["clr"]), cdc) == golden["deformable_between"][1]
    for p, want, hit in zip(polys, golden["path_free"], golden["path_free_hit"]):
        r, h = om.path_collision_free(p, float(golden["clr"]), cdc)
        assert r == want and np.array_equal(h, hit, equal_nan=True)


def test_ellipse_wrapper(golden, orc):
    # BaseMap::fillRandomStateInEllipse (src/base_map.cpp:29-43): dist*ratio and planar override
    q = golden["quad"]
    got = orc.ellipsoid_point(q[0], q[1], np.sqrt(((q[1] - q[0]) ** 2).sum()) * 1.2, float(golden["ell_u"]), golden["ell_g"])
    assert np.allclose(got, golden["ell"], rtol=0, atol=1e-12)
    assert golden["ell_planar"][2] == q[0][2]
    assert np.array_equal(golden["ell_planar"][:2], golden["ell"][:2])
