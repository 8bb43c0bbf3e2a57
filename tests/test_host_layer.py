"""The C++ class surface (include/ctopprm/*.hpp: ESDFMap, BaseMap, TopologicalPRMClustering) on top of the
C ABI.  CPU part: the library loads, exports its entry points, and the YAML subset reader handles the
planner's schema.  GPU part: the classes give the reference's results (golden vectors from the
reference's own code, and the oracle)."""
import os
import tempfile

import numpy as np
import pytest

from ctopprm_learn_b200 import hostlib, synth

CLR, CDC = synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK


def test_host_library_loads_and_exports():
    L = hostlib.lib()
    for name in ("ctph_map_load", "ctph_get_clearence", "ctph_edge_nodes", "ctph_deformable", "ctph_sample_path",
                 "ctph_planner_new", "ctph_planner_sample_multiple", "ctph_planner_graph", "ctph_shorten_path",
                 "ctph_remove_equivalent", "ctph_find_geometrical_paths"):
        assert hasattr(L, name)


def test_yaml_subset_reader():
    y = hostlib.DEFAULT_YAML.format(map="maps/x.npy", clr=0.2, cdc=0.05, planar="false", n=1000, s=[1.0, 2.5, -3.0],
                                    g=[4, 5, 6])
    assert hostlib.yaml_get(y, "min_clearance") == "0.2"
    assert hostlib.yaml_get(y, "map") == "maps/x.npy"
    assert hostlib.yaml_get(y, "start.position") == "[1.0, 2.5, -3.0]"
    assert hostlib.yaml_get(y, "end.position") == "[4, 5, 6]"
    assert hostlib.yaml_get(y, "cutoof_distance_ratio_to_shortest") == "1.3"
    assert hostlib.yaml_get(y, "nope") is None and hostlib.yaml_get(y, "start.nope") is None
    assert hostlib.yaml_get("a: 1 # comment\nb:\n  c:\n    d: 'q'\ne: 2\n", "b.c.d") == "'q'"


def test_missing_map_file_is_reported():
    with pytest.raises(RuntimeError):
        hostlib.HostMap("/nonexistent/map.npy")


@pytest.fixture(scope="module")
def gold_host(golden):
    fn = tempfile.mktemp(suffix=".npy")
    synth.write_map(fn, golden["vox"], golden["centroid"], golden["extents"])
    m = hostlib.HostMap(fn)
    yield m
    m.close()
    os.unlink(fn)


@pytest.mark.gpu
def test_esdf_map_scalar_surface(golden, gold_host):
    m = gold_host
    lo, hi = m.min_max()
    assert np.array_equal(lo, golden["min_pos"]) and np.array_equal(hi, golden["max_pos"])
    for i in list(range(0, 4000, 97)) + list(range(4000, 4068)):
        p = golden["pts"][i]
        got = m.get_clearence(p)
        want = golden["clearance"][i]
        assert (np.isnan(got) and np.isnan(want)) or got == want
        assert m.in_collision(p, CLR) == bool(golden["in_collision"][i])
    for i in range(0, 1500, 61):
        g, c = m.gradient(golden["pts"][i])
        assert np.array_equal(g, golden["grad"][i], equal_nan=True) and np.array_equal(c, golden["centre"][i], equal_nan=True)


@pytest.mark.gpu
def test_base_map_segment_surface(golden, gold_host):
    m = gold_host
    for i in list(range(12)) + list(range(12, 3000, 59)):
        fr, hit = m.edge_nodes(golden["a"][i], golden["b"][i], CLR, CDC)
        assert fr == bool(golden["free_nodes"][i])
        assert np.array_equal(hit, golden["hit"][i], equal_nan=True)
        assert m.edge_guards(golden["a"][i], golden["b"][i], CLR, CDC) == bool(golden["free_guards"][i])


def _split(g, key_pts, key_off):
    off = g[key_off]
    return [g[key_pts][off[i]:off[i + 1]] for i in range(len(off) - 1)]


@pytest.mark.gpu
def test_base_map_polyline_surface(golden, gold_host):
    m = gold_host
    polys = _split(golden, "poly_pts", "poly_off")
    lens, nums = golden["poly_len"], golden["poly_num"]
    pos = 0
    for p, L, k in zip(polys, lens, nums):
        assert np.array_equal(m.sample_path(p, L, k), golden["sampled"][pos:pos + int(k)])
        pos += int(k)
    for i, j, want in zip(golden["pair_i"], golden["pair_j"], golden["deformable"]):
        k = float(np.ceil(max(lens[i], lens[j]) / CDC) + 1)
        assert m.deformable(polys[i], lens[i], polys[j], lens[j], k, 0.02, CDC) == want
    q = golden["quad"]
    assert m.deformable_between(q[0], q[1], q[2], q[3], -10.0, CDC) == golden["deformable_between"][0]
    assert m.deformable_between(q[0], q[1], q[2], q[3], CLR, CDC) == golden["deformable_between"][1]
    for p, want, hit in zip(polys, golden["path_free"], golden["path_free_hit"]):
        r, h = m.path_collision_free(p, CLR, CDC)
        assert r == bool(want) and np.array_equal(h, hit, equal_nan=True)


@pytest.fixture(scope="module")
def c1_host(orc, c1_map):
    vox, c, e = c1_map
    fn = tempfile.mktemp(suffix=".npy")
    synth.write_map(fn, vox, c, e)
    m = hostlib.HostMap(fn)
    yield m, orc.OracleMap(vox, c, e), c, e, fn
    m.close()
    os.unlink(fn)


def _yaml(fn, n, s, g, planar=False):
    return hostlib.DEFAULT_YAML.format(map=fn, clr=CLR, cdc=CDC, planar="true" if planar else "false", n=n, s=list(s), g=list(g))


@pytest.mark.gpu
def test_planner_sample_multiple_from_shared_sample_file(c1_host, orc):
    m, om, c, e, fn = c1_host
    n = 1200
    s, g = synth.default_query(c, e)
    cand = synth.ellipsoid_candidates(s, g, 5 * n, 17)
    pl = hostlib.HostPlanner(m, _yaml(fn, n, s, g), s, g)
    pl.set_candidates(cand)
    pl.sample_multiple(n)
    nodes, row_ptr, col, w = pl.graph()
    want_nodes = om.sample_reject(s, g, cand, CLR, n)[0]
    assert np.array_equal(nodes, want_nodes)
    want = om.build_prm(want_nodes, CLR, CDC, nthreads=8)
    assert np.array_equal(row_ptr, want["row_ptr"]) and np.array_equal(col, want["col"]) and np.array_equal(w, want["w"])
    ids, length = pl.shortest()
    assert ids[0] == 0 and ids[-1] == 1
    # Dijkstra over the re-materialised HeapNode graph: same length as over the oracle's CSR
    import heapq
    dist = np.full(n, np.inf)
    dist[0] = 0
    pq = [(0.0, 0)]
    while pq:
        d, u = heapq.heappop(pq)
        if d > dist[u]:
            continue
        for t in range(want["row_ptr"][u], want["row_ptr"][u + 1]):
            v = want["col"][t]
            if d + want["w"][t] < dist[v]:
                dist[v] = d + want["w"][t]
                heapq.heappush(pq, (dist[v], v))
    assert abs(length - dist[1]) <= 1e-5 * dist[1]
    # findShortestPath ran on the device (ctp_sssp): identical to the oracle's Dijkstra over the same CSR
    d0, p0 = orc.sssp(want["row_ptr"], want["col"], want["w"], [0])
    assert length == d0[1]
    chain = [1]
    while chain[-1] != 0:
        chain.append(int(p0[chain[-1]]))
    assert np.array_equal(ids, chain[::-1])
    # roadmap CSV from the pointer graph and straight from the CSR: same node rows, same set of edge rows
    import tempfile as _tf
    with _tf.TemporaryDirectory() as d_:
        pl.save_roadmaps(d_ + "/dense.csv", d_ + "/csr.csv")
        a_rows = open(d_ + "/dense.csv").read().strip().split("\n")
        b_rows = open(d_ + "/csr.csv").read().strip().split("\n")
        assert a_rows[:n] == b_rows[:n] and len(a_rows) == len(b_rows) == n + len(col)
        assert sorted(a_rows[n:]) == sorted(b_rows[n:])
    # multi-source flood (core of wavefrontFill) through the class
    dist, pred, cluster = pl.flood([0, 1, 500])
    dm, pm, lm = orc.sssp(want["row_ptr"], want["col"], want["w"], [0, 1, 500], want_label=True)
    assert np.array_equal(dist, dm) and np.array_equal(pred, pm) and np.array_equal(cluster, lm)
    # shorten_path through the class (static, uses the planner's statics) against the oracle
    path = nodes[ids]
    plen = orc.path_length(path)
    for forward in (True, False):
        got_pts, got_len = hostlib.shorten_path(path, plen, forward)
        w_pts, _, w_len, w_st, _ = om.shorten_path(path, plen, forward, CLR, CDC, cap=512)
        assert np.array_equal(got_pts, w_pts) and (got_len == w_len or w_st == 1)
    pl.close()


@pytest.mark.gpu
def test_planner_device_sampling_and_entry_point(c1_host):
    m, om, c, e, fn = c1_host
    n = 800
    s, g = synth.default_query(c, e)
    pl = hostlib.HostPlanner(m, _yaml(fn, n, s, g), s, g)
    pl.sample_multiple(n)  # candidates generated on the device (seeded Philox)
    nodes, row_ptr, col, w = pl.graph()
    assert np.array_equal(nodes[0], s) and np.array_equal(nodes[1], g)
    clear = om.clearance(nodes[2:])
    assert (clear >= CLR).all()  # every sampled node passed the rejection test
    want = om.build_prm(nodes, CLR, CDC, nthreads=8)
    assert np.array_equal(col, want["col"]) and np.array_equal(w, want["w"])
    pl.close()
    with tempfile.TemporaryDirectory() as d:
        lens = m.find_geometrical_paths(_yaml(fn, n, s, g), s, g, d + "/")
        assert len(lens) >= 1 and np.isfinite(lens).all() and lens[0] >= np.linalg.norm(g - s) - 1e-9
        rows = open(os.path.join(d, "roadmap_all0.csv")).read().strip().split("\n")
        assert len(rows[0].split(",")) == 3 and len(rows[-1].split(",")) == 6  # nodes then edges
        assert sum(1 for r in rows if len(r.split(",")) == 3) == n


@pytest.mark.gpu
def test_remove_equivalent_and_too_long(c1_host, orc):
    m, om, c, e, fn = c1_host
    s, g = synth.default_query(c, e)
    pl = hostlib.HostPlanner(m, _yaml(fn, 100, s, g), s, g)  # sets the planner statics (clearance, cdc, map)
    rng = np.random.default_rng(4)
    base = np.stack([s + (g - s) * t for t in np.linspace(0, 1, 6)])
    paths = []
    for i in range(6):
        p = base.copy()
        p[1:-1] += rng.standard_normal((4, 3)) * (0.05 if i < 3 else 1.5)
        paths.append(p)
    lens = [orc.path_length(p) for p in paths]
    got = hostlib.remove_equivalent(paths, lens)
    want = om.remove_equivalent(paths, lens, CLR, CDC)
    assert np.array_equal(got, want)
    keep = hostlib.remove_too_long([1.0, 1.2, 1.31, 5.0, 1.29])
    assert list(keep) == [0, 1, 4]
    pl.close()


@pytest.mark.gpu
def test_fill_random_state_in_ellipse(c1_host):
    m, om, c, e, fn = c1_host
    s, g = synth.default_query(c, e)
    a = m.fill_random_state_in_ellipse(s, g, 1.2, False, 99, 300)
    b = m.fill_random_state_in_ellipse(s, g, 1.2, False, 99, 300)
    assert np.array_equal(a, b) and len(np.unique(a, axis=0)) == 300
    two_major = 1.2 * np.linalg.norm(g - s)
    assert (np.linalg.norm(a - s, axis=1) + np.linalg.norm(a - g, axis=1) <= two_major * (1 + 1e-12)).all()
    p = m.fill_random_state_in_ellipse(s, g, 1.2, True, 5, 10)
    assert (p[:, 2] == s[2]).all()


@pytest.mark.gpu
def test_example_main_writes_the_reference_csv_outputs(c1_host, tmp_path):
    """examples/ctopprm_main.cpp: YAML -> ESDFMap::load -> find_geometrical_paths -> the CSVs of src/main.cpp:269-305"""
    import subprocess
    from conftest import ROOT
    m, om, c, e, fn = c1_host
    exe = str(tmp_path / "ctopprm_main")
    libdir = os.path.join(ROOT, "ctopprm_learn_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "ctopprm_main.cpp"), "-L" + libdir, "-lctopprm_host", "-lctopprm_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(open(os.path.join(ROOT, "configs", "default.yaml")).read().replace("maps/columns_c1.npy", fn))
    out = str(tmp_path) + "/out_"
    # the reference names its metric files <output_folder><config path>method_...csv (src/main.cpp:269-271)
    r = subprocess.run([exe, "--config", "cfg.yaml", "--output_folder", out], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert "number of paths found" in r.stdout
    tag = out + "cfg.yaml"
    num = float(open(tag + "method_clustering_num.csv").read().split()[-1])
    secs = float(open(tag + "method_clustering_time.csv").read().split()[-1])
    assert num >= 1 and secs > 0
    lens = [float(x) for x in open(tag + "method_clustering_length.csv").read().split()]
    assert len(lens) == int(num) and lens == sorted(lens)
    rows = open(out + "roadmap_clustering_roadmap_shortened_unique_path0_0.csv").read().strip().split("\n")
    assert all(len(r_.split(",")) == 6 for r_ in rows)
    first, last = [float(x) for x in rows[0].split(",")[:3]], [float(x) for x in rows[-1].split(",")[3:]]
    s, g = synth.default_query(c, e)
    assert np.allclose(first, s, atol=1e-3) and np.allclose(last, g, atol=1e-3)  # 6 significant digits in the CSV
    # sum of the printed segment lengths ~ reported length (the reference's own invariant, :2463-2473)
    seg = np.array([[float(x) for x in r_.split(",")] for r_ in rows])
    assert abs(np.sqrt(((seg[:, 3:] - seg[:, :3]) ** 2).sum(1)).sum() - lens[0]) < 1e-3
    # a missing map file ends the program the way cnpy does in the reference (uncaught std::runtime_error)
    r2 = subprocess.run([exe, "--config", "cfg.yaml", "--output_folder", out, "--map", "/nonexistent.npy"], capture_output=True,
                        text=True, cwd=str(tmp_path))
    assert r2.returncode != 0 and "Unable to open file" in r2.stderr


@pytest.mark.gpu
def test_homotopy_decision_and_seed_test(c1_host, orc):
    """the tail of isNewHomotopyClass (:1702-1726) and the seed visibility test of clusterGraph (:1856-1888),
    batched through the C++ layer, against the same composition of oracle functions"""
    m, om, c, e, fn = c1_host
    n = 900
    s, g = synth.default_query(c, e)
    cand = synth.ellipsoid_candidates(s, g, 5 * n, 23)
    pl = hostlib.HostPlanner(m, _yaml(fn, n, s, g), s, g)
    pl.set_candidates(cand)
    pl.sample_multiple(n)
    nodes, row_ptr, col, w = pl.graph()
    # path pairs: shortest paths via two different waypoints (min/max connections of a cluster pair look like this)
    d0, p0 = orc.sssp(row_ptr, col, w, [0])
    d1, p1 = orc.sssp(row_ptr, col, w, [1])

    def via(v):
        a = [v]
        while a[-1] != 0:
            a.append(int(p0[a[-1]]))
        b = [v]
        while b[-1] != 1:
            b.append(int(p1[b[-1]]))
        return nodes[a[::-1] + b[1:]]
    reach = np.flatnonzero(np.isfinite(d0) & np.isfinite(d1))
    rng = np.random.default_rng(6)
    wps = rng.choice(reach[2:], 8, replace=False)
    mins = [via(int(v)) for v in wps[:4]]
    maxs = [via(int(v)) for v in wps[4:]]
    ml = [orc.path_length(p) for p in mins]
    xl = [orc.path_length(p) for p in maxs]
    got = hostlib.new_homotopy_batch(mins, maxs, ml, xl)
    for i in range(4):
        a, _, la, sa, _ = om.shorten_path(mins[i], ml[i], False, CLR, CDC, cap=1024)
        a, _, la, sa2, _ = om.shorten_path(a, la, True, CLR, CDC, cap=1024)
        b, _, lb, sb, _ = om.shorten_path(maxs[i], xl[i], True, CLR, CDC, cap=1024)
        b, _, lb, sb2, _ = om.shorten_path(b, lb, False, CLR, CDC, cap=1024)
        k = float(np.ceil(max(la, lb) / CDC) + 1)
        assert got[i] == om.deformable(a, la, b, lb, k, CLR, CDC)
    # seed test: candidate accepted iff no seed is a neighbour / visible and there are >= 2 seeds
    adj = {(r, cc) for r in range(n) for cc in col[row_ptr[r]:row_ptr[r + 1]]}
    seeds = [int(x) for x in rng.choice(reach[2:], 3, replace=False)]
    for cand_id in rng.choice(reach[2:], 30, replace=False):
        cand_id = int(cand_id)
        if cand_id in seeds:
            continue
        want = True
        blocked = 0
        for sd in seeds:
            if (sd, cand_id) in adj:
                want = False
                break
            if om.edge_nodes(nodes[sd][None], nodes[cand_id][None], CLR, CDC)[0][0]:
                want = False
                break
            blocked += 1
        want = want and blocked >= 2
        assert pl.seed_accepted(cand_id, seeds) == want
    pl.close()


@pytest.mark.gpu
def test_map_from_occupancy_saved_in_the_map_file_format(orc, tmp_path):
    """ESDFMap::loadFromOccupancy + save: the device distance transform written as the four NPY records of
    map.py:124-132; numpy reads the file back, the loader accepts it, the field is the oracle's"""
    n = 48
    occ = synth.blocks_occupancy(n, 10, 31)
    e = np.array([4.8] * 3)
    c = np.array([0.013, -0.007, e[2] / 2 + 0.011])
    m = hostlib.HostMap.from_occupancy(occ, c, e)
    fn = str(tmp_path / "edt_map.npy")
    m.save(fn)
    vox, c2, e2, res = synth.read_map(fn)
    assert vox.dtype == np.float32 and vox.shape == (n, n, n) and int(res) == n
    assert np.array_equal(c2, c) and np.array_equal(e2, e)
    assert np.array_equal(vox, orc.edt_signed(occ, e[0] / n)[0])
    m2 = hostlib.HostMap(fn)
    om = orc.OracleMap(vox, c, e)
    pts = c + (np.random.default_rng(3).random((64, 3)) - 0.5) * e
    want = om.clearance(pts)
    for p, w in zip(pts, want):
        assert m.get_clearence(p) == w or (np.isnan(w) and np.isnan(m.get_clearence(p)))
        assert m2.get_clearence(p) == w or (np.isnan(w) and np.isnan(m2.get_clearence(p)))
    m.close()
    m2.close()
