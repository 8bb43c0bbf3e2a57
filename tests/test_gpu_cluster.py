"""The floods of the cluster stage (ctp_cluster_flood: wavefrontFill's multi-source flood and addCentroid's re-flood,
include/topological_prm_clustering.hpp:826-962, :1276-1374) on the device against the oracle's sequential loop:
distances, predecessors, labels, the connection records IN THE ORDER the loop appends them, and the update counts,
all bit-exact."""
import numpy as np
import pytest

from ctopprm_learn_b200 import capi, synth

pytestmark = pytest.mark.gpu

CLR, CDC = synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK
DBL_MAX = np.finfo(np.float64).max


@pytest.fixture(scope="module")
def c1(orc, c1_map):
    vox, c, e = c1_map
    m = capi.DeviceMap(vox, c, e, layouts=capi.LAYOUT_PLAIN | capi.LAYOUT_TEX)
    yield m, orc.OracleMap(vox, c, e), c, e
    m.close()


def _random_graph(n, deg, seed, components=1):
    """symmetric weighted graph as CSR (columns ascending), optionally split into components"""
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 3))
    comp = rng.integers(0, components, n)
    rows = [set() for _ in range(n)]
    for i in range(n):
        same = np.flatnonzero(comp == comp[i])
        for j in rng.choice(same, size=min(deg, len(same)), replace=False):
            if j != i:
                rows[i].add(int(j))
                rows[int(j)].add(i)
    rp = np.zeros(n + 1, dtype=np.int32)
    col, w = [], []
    for i in range(n):
        js = sorted(rows[i])
        rp[i + 1] = rp[i] + len(js)
        col += js
        w += [float(np.linalg.norm(pts[i] - pts[j])) for j in js]
    return rp, np.array(col, dtype=np.int32), np.array(w)


def _check_sequence(orc, m, graphs, seed_lists, extra):
    """mode 0 from seed_lists, then one mode 1 flood per entry of extra (per query: node ids to add, one per round)"""
    q = len(graphs)
    n = len(graphs[0][0]) - 1
    rp = np.stack([g[0] for g in graphs])
    off = np.zeros(q + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(g[1]) for g in graphs])
    col = np.concatenate([g[1] for g in graphs])
    w = np.concatenate([g[2] for g in graphs])
    rm = m.roadmap(rp, col, w, off, max_seeds=8)
    seeds = [list(s) for s in seed_lists]
    state = [None] * q
    dist, prev, label, recs, upd = rm.flood(seeds, mode=0)
    rounds = max(len(x) for x in extra)
    capped = [False] * q
    n_capped = 0
    for r in range(rounds + 1):
        active = [True] * q if r == 0 else [r - 1 < len(extra[i]) and not capped[i] for i in range(q)]
        for i in range(q):
            if not active[i]:
                continue
            g = graphs[i]
            if r == 0:
                o = orc.cluster_flood(g[0], g[1], g[2], seeds[i], mode=0)
            else:
                o = orc.cluster_flood(g[0], g[1], g[2], seeds[i], mode=1, dist=state[i][0], prev=state[i][1], cluster=state[i][2])
            if r > 0 and o[5] >= n - 1:
                # the reference stops improving once cnt reaches the node count (:1319): such a re-flood is not
                # reproduced on the device, it is reported through the update count and the planner stops there
                assert int(upd[i]) >= n - 1
                capped[i] = True
                n_capped += 1
                continue
            state[i] = o[:3]
            od = np.where(o[0] == DBL_MAX, np.inf, o[0])
            assert np.array_equal(dist[i], od), f"round {r} query {i}: distances"
            assert np.array_equal(label[i], o[2]), f"round {r} query {i}: labels"
            assert np.array_equal(prev[i], o[1]), f"round {r} query {i}: predecessors"
            assert len(recs[i][1]) == len(o[4]), f"round {r} query {i}: record count {len(recs[i][1])} != {len(o[4])}"
            assert np.array_equal(recs[i][0], o[3]), f"round {r} query {i}: record nodes / order"
            assert np.array_equal(recs[i][1], o[4]), f"round {r} query {i}: record distances"
            assert np.array_equal(recs[i][2], o[6]), f"round {r} query {i}: record labels"
            assert int(upd[i]) == o[5], f"round {r} query {i}: update count {int(upd[i])} != {o[5]}"
        if r == rounds:
            break
        nxt = []
        for i in range(q):
            if r < len(extra[i]) and not capped[i]:
                seeds[i] = seeds[i] + [extra[i][r]]
                nxt.append(seeds[i])
            else:
                nxt.append([])  # inactive: nothing of this query may change
        before = (dist.copy(), prev.copy(), label.copy())
        dist, prev, label, recs, upd = rm.flood(nxt, mode=1)
        for i in range(q):
            if not nxt[i]:
                assert np.array_equal(dist[i], before[0][i]) and np.array_equal(prev[i], before[1][i])
                assert np.array_equal(label[i], before[2][i]) and len(recs[i][1]) == 0
    rm.close()
    return n_capped


@pytest.mark.parametrize("n,deg,comps,seed", [(300, 3, 1, 1), (2000, 4, 1, 2), (5000, 6, 3, 3), (40, 2, 2, 4), (20000, 5, 1, 5)])
def test_floods_on_synthetic_graphs(c1, orc, n, deg, comps, seed):
    m = c1[0]
    rng = np.random.default_rng(seed + 100)
    graphs = [_random_graph(n, deg, seed * 10 + i, comps) for i in range(3)]
    seed_lists = [[0, 1], [0, 1, int(rng.integers(2, n))], list(rng.choice(n, 4, replace=False))]
    extra = [list(rng.choice(np.arange(2, n), 3, replace=False)), [int(rng.integers(2, n))], []]
    extra = [[int(x) for x in e if x not in seed_lists[i]] for i, e in enumerate(extra)]
    _check_sequence(orc, m, graphs, seed_lists, extra)


@pytest.mark.parametrize("n", [600, 3000])
def test_floods_on_roadmaps(c1, orc, n):
    m, om, c, e = c1
    q = 4
    s, g = synth.random_queries(c, e, q, 31)
    cand = np.stack([synth.ellipsoid_candidates(s[i], g[i], 6 * n, 3100 + i) for i in range(q)])
    out = m.sample_multiple(s, g, cand, n, CLR, CDC)
    graphs = []
    for i in range(q):
        a, b = out["nnz_off"][i], out["nnz_off"][i + 1]
        graphs.append((out["row_ptr"][i], out["col"][a:b], out["w"][a:b]))
    rng = np.random.default_rng(n)
    seed_lists = [[0, 1]] * q
    extra = [[int(x) for x in rng.choice(np.arange(2, n), 4, replace=False)] for _ in range(q)]
    _check_sequence(orc, m, graphs, seed_lists, extra)


def test_flood_argument_errors(c1):
    m = c1[0]
    rp, col, w = _random_graph(50, 3, 9)
    off = np.array([0, len(col)], dtype=np.int64)
    with pytest.raises(capi.CtpError):
        m.roadmap(rp, col, w, off, max_seeds=0)
    bad = col.copy()
    bad[0] = 50
    with pytest.raises(capi.CtpError):
        m.roadmap(rp, bad, w, off)
    rm = m.roadmap(rp, col, w, off, max_seeds=4)
    with pytest.raises(capi.CtpError):
        rm.flood([[0, 50]])
    with pytest.raises(capi.CtpError):
        rm.flood([[0, 1]], mode=2)
    rm.close()


# ---- the whole cluster stage through the C++ class surface (csrc/host/cluster_graph.cpp) -------------------------------
import os
import tempfile

from ctopprm_learn_b200 import hostlib


@pytest.fixture(scope="module")
def c1_host(orc, c1_map):
    vox, c, e = c1_map
    fn = tempfile.mktemp(suffix=".npy")
    synth.write_map(fn, vox, c, e)
    m = hostlib.HostMap(fn)
    yield m, orc.OracleMap(vox, c, e), c, e, fn
    m.close()
    os.unlink(fn)


def _yaml(fn, n, s, g, **kw):
    y = hostlib.DEFAULT_YAML.format(map=fn, clr=CLR, cdc=CDC, planar="false", n=n, s=list(s), g=list(g))
    for k, v in kw.items():
        lines = [f"{k}: {v}" if ln.startswith(k + ":") else ln for ln in y.split("\n")]
        y = "\n".join(lines)
    return y


def _same_paths(got, want):
    return len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))


@pytest.mark.parametrize("n,qseed,num_clusters", [(600, 5, 2), (1500, 6, 2), (1500, 7, 4), (4000, 8, 2)])
def test_cluster_graph_vs_oracle(c1_host, orc, n, qseed, num_clusters):
    """clusterGraph (:1832-1952): stitched cluster paths, their DFS lengths, final seeds and labels"""
    m, om, c, e, fn = c1_host
    s, g = synth.random_queries(c, e, 1, qseed, cfg="C1")
    s, g = s[0], g[0]
    cand = synth.ellipsoid_candidates(s, g, 6 * n, 900 + qseed)
    pl = hostlib.HostPlanner(m, _yaml(fn, n, s, g, num_clusters=num_clusters), s, g)
    pl.set_candidates(cand)
    pl.sample_multiple(n)
    nodes0 = om.sample_reject(s, g, cand, CLR, n)[0]
    want = om.build_prm(nodes0, CLR, CDC, nthreads=8)
    d0, _ = orc.sssp(want["row_ptr"], want["col"], want["w"], [0])
    if not np.isfinite(d0[1]):
        pytest.skip("goal not reachable on this roadmap")
    seed_cand = [int(x) for x in np.random.default_rng(qseed).integers(2, n, 64)]
    w_paths, w_len, w_seeds, w_labels = orc.cluster_graph(om, nodes0, want["row_ptr"], want["col"], want["w"], CLR, CDC,
                                                          d0[1] * 1.5, num_clusters=num_clusters, seed_candidates=seed_cand)
    capped = orc.cluster_capped_refloods()
    paths, lens, seeds, labels, stats = pl.cluster_graph(d0[1] * 1.5, seed_cand)
    pl.close()
    assert stats["floods"] == 1 and stats["seeds"] == len(seeds)
    if capped:  # the node-count cap of addCentroid fired in the sequential loop: declared not reproduced
        assert stats["capped_refloods"] >= 1
        return
    assert stats["capped_refloods"] == 0
    assert np.array_equal(seeds, w_seeds)
    assert np.array_equal(labels, w_labels)
    assert np.array_equal(lens, w_len)
    assert _same_paths(paths, w_paths)
    assert len(seeds) >= 2 and len(paths) >= 1


@pytest.mark.parametrize("n,qseed", [(800, 11), (2500, 12)])
def test_find_geometrical_paths_vs_oracle(c1_host, orc, n, qseed):
    """the whole entry point (:2489-2684) for one start-goal pair: the distinct paths that come out, point for point"""
    m, om, c, e, fn = c1_host
    s, g = synth.random_queries(c, e, 1, qseed, cfg="C1")
    s, g = s[0], g[0]
    cand = synth.ellipsoid_candidates(s, g, 6 * n, 700 + qseed)
    want = orc.find_geometrical_paths(om, s, g, cand, n, CLR, CDC, nthreads=8)
    lens, paths = m.find_geometrical_paths(_yaml(fn, n, s, g), s, g, "", cand=cand, want_paths=True)
    if want["capped_refloods"]:
        return
    assert np.array_equal(lens, want["lengths"])
    assert _same_paths(paths, want["paths"])
    assert len(lens) >= 1 and (np.diff(lens) >= 0).all()
