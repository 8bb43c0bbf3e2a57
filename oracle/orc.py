"""ctypes binding of the test oracle (oracle/ctp_oracle.c) and, when built, of the
reference's own code (oracle/_ref/libctopprm_ref.so).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by
the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "_build", "libctp_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libctopprm_ref.so")


def build(ref_root="/root/reference"):
    """Compile the oracle (always) and the reference shim build (only where the reference
    tree is mounted); outputs stay under oracle/_build and oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    if os.path.isdir(os.path.join(ref_root, "src")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref", f"REF={ref_root}"])


class _Map(C.Structure):
    _fields_ = [("vox", C.c_void_p), ("n", C.c_int), ("N", C.c_double), ("c", C.c_double * 3),
                ("e", C.c_double * 3), ("max_e", C.c_double)]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_lib = None


def _L():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _lib = C.CDLL(ORACLE_SO)
        _lib.orc_clearance.restype = C.c_double
        _lib.orc_interpolate.restype = C.c_double
        _lib.orc_path_length.restype = C.c_double
        for f in ("orc_edge_nodes_batch", "orc_edge_guards_batch", "orc_edge_index_batch", "orc_build_prm"):
            getattr(_lib, f).restype = C.c_int64
    return _lib


class OracleMap:
    def __init__(self, vox, centroid, extents):
        self.vox = np.ascontiguousarray(vox, dtype=np.float32)
        self.n = self.vox.shape[0]
        self.m = _Map()
        c = _f64(centroid)
        e = _f64(extents)
        _L().orc_map_init(C.byref(self.m), _p(self.vox), C.c_int(self.n), _p(c), _p(e))
        self.centroid, self.extents = c, e

    def min_pos(self):
        o = np.empty(3)
        _L().orc_min_pos(C.byref(self.m), _p(o))
        return o

    def max_pos(self):
        o = np.empty(3)
        _L().orc_max_pos(C.byref(self.m), _p(o))
        return o

    def interpolate(self, q):
        q = _f64(q)
        return _L().orc_interpolate(C.byref(self.m), _p(q))

    def clearance(self, xyz, nthreads=1):
        xyz = _f64(xyz).reshape(-1, 3)
        out = np.empty(len(xyz))
        _L().orc_clearance_batch(C.byref(self.m), _p(xyz), C.c_int64(len(xyz)), _p(out), C.c_int(nthreads))
        return out

    def gradient(self, xyz, nthreads=1):
        xyz = _f64(xyz).reshape(-1, 3)
        g = np.empty_like(xyz)
        c = np.empty_like(xyz)
        _L().orc_gradient_batch(C.byref(self.m), _p(xyz), C.c_int64(len(xyz)), _p(g), _p(c), C.c_int(nthreads))
        return g, c

    def in_collision(self, xyz, clr):
        c = self.clearance(xyz)
        return ~np.isfinite(c) | (c < clr)

    def edge_nodes(self, a, b, clr, cdc, nthreads=1, full=True):
        a = _f64(a).reshape(-1, 3)
        b = _f64(b).reshape(-1, 3)
        m = len(a)
        free = np.empty(m, dtype=np.uint8)
        hit = np.empty((m, 3)) if full else None
        hit_idx = np.empty(m, dtype=np.int32) if full else None
        lookups = np.empty(m, dtype=np.int32) if full else None
        tot = _L().orc_edge_nodes_batch(C.byref(self.m), _p(a), _p(b), C.c_int64(m), C.c_double(clr), C.c_double(cdc),
                                        _p(free), _p(hit), _p(hit_idx), _p(lookups), C.c_int(nthreads))
        return free.astype(bool), hit, hit_idx, lookups, int(tot)

    def edge_guards(self, a, b, clr, cdc, nthreads=1):
        a = _f64(a).reshape(-1, 3)
        b = _f64(b).reshape(-1, 3)
        free = np.empty(len(a), dtype=np.uint8)
        tot = _L().orc_edge_guards_batch(C.byref(self.m), _p(a), _p(b), C.c_int64(len(a)), C.c_double(clr),
                                         C.c_double(cdc), _p(free), C.c_int(nthreads))
        return free.astype(bool), int(tot)

    def edge_index(self, nodes, frm, to, clr, cdc, nthreads=1):
        nodes = _f64(nodes).reshape(-1, 3)
        frm = np.ascontiguousarray(frm, dtype=np.int32)
        to = np.ascontiguousarray(to, dtype=np.int32)
        free = np.empty(len(frm), dtype=np.uint8)
        hit_idx = np.empty(len(frm), dtype=np.int32)
        tot = _L().orc_edge_index_batch(C.byref(self.m), _p(nodes), _p(frm), _p(to), C.c_int64(len(frm)),
                                        C.c_double(clr), C.c_double(cdc), _p(free), _p(hit_idx), C.c_int(nthreads))
        return free.astype(bool), hit_idx, int(tot)

    def deformable(self, p1, len1, p2, len2, num_check, clr, cdc):
        p1 = _f64(p1).reshape(-1, 3)
        p2 = _f64(p2).reshape(-1, 3)
        return _L().orc_deformable(C.byref(self.m), _p(p1), C.c_int(len(p1)), C.c_double(len1), _p(p2),
                                   C.c_int(len(p2)), C.c_double(len2), C.c_double(num_check), C.c_double(clr),
                                   C.c_double(cdc))

    def deformable_between(self, s, e, b1, b2, clr, cdc):
        s, e, b1, b2 = (_f64(x) for x in (s, e, b1, b2))
        return _L().orc_deformable_between(C.byref(self.m), _p(s), _p(e), _p(b1), _p(b2), C.c_double(clr), C.c_double(cdc))

    def path_collision_free(self, pts, clr, cdc):
        pts = _f64(pts).reshape(-1, 3)
        hit = np.empty(3)
        r = _L().orc_path_collision_free(C.byref(self.m), _p(pts), C.c_int(len(pts)), C.c_double(clr), C.c_double(cdc), _p(hit))
        return r, hit

    def sample_reject(self, start, goal, cand, clr, n):
        cand = _f64(cand).reshape(-1, 3)
        s = _f64(start)
        g = _f64(goal)
        nodes = np.zeros((n, 3))
        accept = np.zeros(len(cand), dtype=np.uint8)
        used = C.c_int64(0)
        filled = _L().orc_sample_reject(C.byref(self.m), _p(s), _p(g), _p(cand), C.c_int64(len(cand)), C.c_double(clr),
                                        C.c_int(n), _p(nodes), _p(accept), C.byref(used))
        return nodes, filled, used.value, accept.astype(bool)

    def build_prm(self, nodes, clr, cdc, k=13, nthreads=1):
        nodes = _f64(nodes).reshape(-1, 3)
        n = len(nodes)
        nbr = np.empty((n, k), dtype=np.int32)
        dfree = np.empty((n, k), dtype=np.uint8)
        row_ptr = np.empty(n + 1, dtype=np.int32)
        col = np.empty(2 * n * k, dtype=np.int32)
        w = np.empty(2 * n * k)
        lookups = C.c_int64(0)
        nnz = _L().orc_build_prm(C.byref(self.m), _p(nodes), C.c_int(n), C.c_int(k), C.c_double(clr), C.c_double(cdc),
                                 _p(nbr), _p(dfree), _p(row_ptr), _p(col), _p(w), C.byref(lookups), C.c_int(nthreads))
        return dict(nbr=nbr, directed_free=dfree.astype(bool), row_ptr=row_ptr, col=col[:nnz], w=w[:nnz],
                    lookups=lookups.value)

    def shorten_path(self, pts, length, forward, clr, cdc, cap=None):
        pts = _f64(pts).reshape(-1, 3)
        if cap is None:
            cap = len(pts) + 64
        out = np.empty((cap, 3))
        origin = np.empty(cap, dtype=np.int32)
        out_len = C.c_double(0)
        status = C.c_int(0)
        lookups = C.c_int64(0)
        n = _L().orc_shorten_path(C.byref(self.m), _p(pts), C.c_int(len(pts)), C.c_double(length), C.c_int(int(forward)),
                                  C.c_double(clr), C.c_double(cdc), _p(out), _p(origin), C.c_int(cap),
                                  C.byref(out_len), C.byref(status), C.byref(lookups))
        return out[:n].copy(), origin[:n].copy(), out_len.value, status.value, lookups.value

    def remove_equivalent(self, paths, lengths, clr, cdc):
        off = np.zeros(len(paths) + 1, dtype=np.int32)
        for i, p in enumerate(paths):
            off[i + 1] = off[i] + len(p)
        pts = _f64(np.concatenate([np.asarray(p).reshape(-1, 3) for p in paths]))
        lengths = _f64(lengths)
        keep = np.empty(len(paths), dtype=np.int32)
        cnt = _L().orc_remove_equivalent(C.byref(self.m), _p(pts), _p(off), _p(lengths), C.c_int(len(paths)),
                                         C.c_double(clr), C.c_double(cdc), _p(keep))
        return keep[:cnt].copy()


def _sample_multiple_batch(fn, handle, start, goal, cand, n, k, clr, cdc, nthreads, full, with_lookups):
    start = _f64(start).reshape(-1, 3)
    goal = _f64(goal).reshape(-1, 3)
    q = len(start)
    cand = _f64(cand).reshape(q, -1, 3)
    M = cand.shape[1]
    n_filled = np.zeros(q, dtype=np.int32)
    n_free = np.zeros(q, dtype=np.int64)
    nnz = np.zeros(q, dtype=np.int64)
    lookups = np.zeros(q, dtype=np.int64)
    out = dict(n_filled=n_filled, n_free=n_free, nnz=nnz)
    bufs = [None] * 4
    if full:
        bufs = [np.empty((q, n, 3)), np.empty((q, n + 1), dtype=np.int32), np.empty((q, 2 * n * k), dtype=np.int32),
                np.empty((q, 2 * n * k))]
        out.update(nodes=bufs[0], row_ptr=bufs[1], col=bufs[2], w=bufs[3])
    args = [handle, _p(start), _p(goal), _p(cand), C.c_int(q), C.c_int64(M), C.c_int(n), C.c_int(k), C.c_double(clr),
            C.c_double(cdc), C.c_int(nthreads), _p(n_filled), _p(n_free), _p(nnz)]
    if with_lookups:
        args.append(_p(lookups))
        out["lookups"] = lookups
    out["short"] = fn(*args, *[_p(b) for b in bufs])
    return out


def sample_multiple_batch(om, start, goal, cand, n, clr, cdc, k=13, nthreads=1, full=False):
    """orc_sample_multiple_batch: whole sampleMultiple for q queries with the restatement."""
    return _sample_multiple_batch(_L().orc_sample_multiple_batch, C.byref(om.m), start, goal, cand, n, k, clr, cdc,
                                  nthreads, full, True)


def sssp(row_ptr, col, w, src, want_label=False):
    """orc_sssp: Dijkstra over one CSR graph from the listed sources -> dist, pred[, label]"""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    w = _f64(w)
    src = np.ascontiguousarray(src, dtype=np.int32).reshape(-1)
    n = len(row_ptr) - 1
    dist = np.empty(n)
    pred = np.empty(n, dtype=np.int32)
    label = np.empty(n, dtype=np.int32) if want_label else None
    _L().orc_sssp(C.c_int(n), _p(row_ptr), _p(col), _p(w), _p(src), C.c_int(len(src)), _p(dist), _p(pred), _p(label))
    return (dist, pred, label) if want_label else (dist, pred)


def cluster_flood(row_ptr, col, w, seeds, mode=0, dist=None, prev=None, cluster=None, cap=None):
    """orc_cluster_flood: one flood of the cluster stage -> dist, prev, cluster, records (nodes m x 2, dist m), updates"""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    w = _f64(w)
    seeds = np.ascontiguousarray(seeds, dtype=np.int32)
    n = len(row_ptr) - 1
    dist = np.empty(n) if dist is None else _f64(dist).copy()
    prev = np.empty(n, dtype=np.int32) if prev is None else np.ascontiguousarray(prev, dtype=np.int32).copy()
    cluster = np.empty(n, dtype=np.int32) if cluster is None else np.ascontiguousarray(cluster, dtype=np.int32).copy()
    cap = len(col) if cap is None else cap
    rn = np.empty((cap, 2), dtype=np.int32)
    rd = np.empty(cap)
    rl = np.empty(cap, dtype=np.int32)
    upd = C.c_int64(0)
    m = _L().orc_cluster_flood(C.c_int(n), _p(row_ptr), _p(col), _p(w), _p(seeds), C.c_int(len(seeds)), C.c_int(mode), _p(dist),
                               _p(prev), _p(cluster), C.c_int(cap), _p(rn), _p(rl), _p(rd), C.byref(upd))
    return dist, prev, cluster, rn[:min(m, cap)].copy(), rd[:min(m, cap)].copy(), upd.value, rl[:min(m, cap)].copy()


def cluster_graph(om, nodes, row_ptr, col, w, clr, cdc, max_path_length, num_clusters=2, max_clusters=10, min_clusters=2,
                  seed_candidates=(), cap_paths=4096, cap_pts=1 << 20):
    """orc_cluster_graph: clusterGraph (:1832-1952) -> (paths [k x 3 arrays], DFS lengths, seeds, labels)"""
    nodes = _f64(nodes).reshape(-1, 3)
    n = len(nodes)
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    w = _f64(w)
    sc = np.ascontiguousarray(seed_candidates, dtype=np.int32)
    pts = np.empty((cap_pts, 3))
    off = np.zeros(cap_paths + 1, dtype=np.int32)
    lens = np.empty(cap_paths)
    seeds = np.full(max(2, max_clusters), -1, dtype=np.int32)
    ns = C.c_int32(0)
    labels = np.empty(n, dtype=np.int32)
    rc = _L().orc_cluster_graph(C.byref(om.m), C.c_int(n), _p(nodes), _p(row_ptr), _p(col), _p(w), C.c_double(clr), C.c_double(cdc),
                                C.c_int(num_clusters), C.c_int(max_clusters), C.c_int(min_clusters), C.c_double(max_path_length),
                                _p(sc), C.c_int(len(sc)), C.c_int(cap_paths), C.c_int(cap_pts), _p(pts), _p(off), _p(lens),
                                _p(seeds), C.byref(ns), _p(labels))
    if rc < 0:
        raise RuntimeError(f"orc_cluster_graph failed: {rc}")
    return [pts[off[i]:off[i + 1]].copy() for i in range(rc)], lens[:rc].copy(), seeds[:ns.value].copy(), labels


def cluster_capped_refloods():
    """re-floods of the last cluster_graph call in which addCentroid's node-count cap (:1319) fired"""
    return int(_L().orc_cluster_capped_refloods())


def find_geometrical_paths(om, start, goal, cand, n, clr, cdc, num_clusters=2, max_clusters=10, min_clusters=2,
                           max_path_length_ratio=1.5, cutoff_ratio=1.3, min_allowed=0.0, seed_candidates=(), nthreads=1):
    """find_geometrical_paths (include/topological_prm_clustering.hpp:2489-2684) for one start-goal pair, composed from
    the restated stages: sampleMultiple -> Dijkstra -> clusterGraph -> shorten both ways -> removeTooLongPaths ->
    removeEquivalentPaths -> shorten both ways -> removeEquivalentPaths -> sort by length.
    -> dict(paths, lengths, cluster_paths, cluster_lengths, seeds, labels, shortest, capped_refloods)"""
    nodes, filled, _, _ = om.sample_reject(start, goal, cand, clr, n)
    if filled < n:
        raise RuntimeError("candidate stream too short")
    g = om.build_prm(nodes, clr, cdc, nthreads=nthreads)
    dist, _ = sssp(g["row_ptr"], g["col"], g["w"], [0])
    out = dict(paths=[], lengths=np.zeros(0), cluster_paths=[], cluster_lengths=np.zeros(0), seeds=np.zeros(0, np.int32),
               labels=None, shortest=float(dist[1]), capped_refloods=0, nodes=nodes, graph=g)
    if not np.isfinite(dist[1]):
        return out
    found, flen, seeds, labels = cluster_graph(om, nodes, g["row_ptr"], g["col"], g["w"], clr, cdc, dist[1] * max_path_length_ratio,
                                               num_clusters, max_clusters, min_clusters, seed_candidates)
    out.update(cluster_paths=found, cluster_lengths=flen, seeds=seeds, labels=labels, capped_refloods=cluster_capped_refloods())

    def both_ways(paths, lens):
        res, rl = [], []
        for p, l in zip(paths, lens):
            for fwd in (False, True):  # shorten_path(path, false) then shorten_path(path) (:2612-2625)
                q, _, ql, status, _ = om.shorten_path(p, l, fwd, clr, cdc, cap=len(p) + 4096)
                if status < 0:
                    raise RuntimeError(f"shorten_path status {status}: the reference would exit")
                if status == 0:
                    p, l = q, ql
            res.append(p)
            rl.append(l)
        return res, rl

    paths, lens = both_ways(found, list(flen))
    keep = remove_too_long(lens, min_allowed, cutoff_ratio)
    paths = [p for p, k in zip(paths, keep) if k]
    lens = [l for l, k in zip(lens, keep) if k]
    for _ in range(2):
        if len(paths) > 1:
            ids = om.remove_equivalent(paths, lens, clr, cdc)
            paths, lens = [paths[i] for i in ids], [lens[i] for i in ids]
        if _ == 0:
            paths, lens = both_ways(paths, lens)
    order = sorted(range(len(paths)), key=lambda i: lens[i])  # std::sort by length (:2656)
    out.update(paths=[paths[i] for i in order], lengths=np.array([lens[i] for i in order]))
    return out


def remove_too_long(lengths, min_allowed, cutoff_ratio):
    lengths = _f64(lengths)
    keep = np.empty(len(lengths), dtype=np.uint8)
    _L().orc_remove_too_long(_p(lengths), C.c_int(len(lengths)), C.c_double(min_allowed), C.c_double(cutoff_ratio), _p(keep))
    return keep.astype(bool)


def edt_signed(occ, pitch):
    """orc_edt_signed: exact signed EDT of an occupancy grid by definition -> (f32 field, signed int32 d2)"""
    occ = np.ascontiguousarray(np.asarray(occ) != 0, dtype=np.uint8)
    n = occ.shape[0]
    vox = np.empty(occ.shape, dtype=np.float32)
    d2 = np.empty(occ.shape, dtype=np.int32)
    _L().orc_edt_signed(_p(occ), C.c_int(n), C.c_double(pitch), _p(vox), _p(d2))
    return vox, d2


def knn_grid(nodes, k=13, nthreads=1):
    nodes = _f64(nodes).reshape(-1, 3)
    n = len(nodes)
    idx = np.empty((n, k), dtype=np.int32)
    d2 = np.empty((n, k), dtype=np.float32)
    _L().orc_knn_grid(_p(nodes), C.c_int(n), C.c_int(k), _p(idx), _p(d2), C.c_int(nthreads))
    return idx, d2


def csr_from_directed(nodes, nbr, directed_free):
    """orc_csr_from_directed: symmetric union of the free directed edges -> row_ptr, col, w"""
    nodes = _f64(nodes).reshape(-1, 3)
    nbr = np.ascontiguousarray(nbr, dtype=np.int32)
    n, k = nbr.shape
    fr = np.ascontiguousarray(directed_free, dtype=np.uint8)
    row_ptr = np.empty(n + 1, dtype=np.int32)
    col = np.empty(2 * n * k, dtype=np.int32)
    w = np.empty(2 * n * k)
    L = _L()
    L.orc_csr_from_directed.restype = C.c_int64
    nnz = L.orc_csr_from_directed(_p(nodes), C.c_int(n), C.c_int(k), _p(nbr), _p(fr), _p(row_ptr), _p(col), _p(w))
    return row_ptr, col[:nnz].copy(), w[:nnz].copy()


def knn(nodes, k=13, nthreads=1):
    nodes = _f64(nodes).reshape(-1, 3)
    n = len(nodes)
    idx = np.empty((n, k), dtype=np.int32)
    d2 = np.empty((n, k), dtype=np.float32)
    _L().orc_knn(_p(nodes), C.c_int(n), C.c_int(k), _p(idx), _p(d2), C.c_int(nthreads))
    return idx, d2


def sample_path(pts, length, num_samples):
    pts = _f64(pts).reshape(-1, 3)
    out = np.empty((int(num_samples), 3))
    rc = _L().orc_sample_path(_p(pts), C.c_int(len(pts)), C.c_double(length), C.c_double(num_samples), _p(out))
    return out, rc


def path_length(pts):
    pts = _f64(pts).reshape(-1, 3)
    return _L().orc_path_length(_p(pts), C.c_int(len(pts)))


def ellipsoid_point(A, B, two_major, u, g):
    A, B, g = _f64(A), _f64(B), _f64(g)
    o = np.empty(3)
    _L().orc_ellipsoid_point(_p(A), _p(B), C.c_double(two_major), C.c_double(u), _p(g), _p(o))
    return o


# ------------------------------ the reference's own code ---------------------------------
class RefMap:
    """The reference's ESDFMap/BaseMap (src/esdf_map.cpp, src/base_map.cpp) compiled against
    shim headers into oracle/_ref.  Loads a map FILE through the reference's own loader."""

    def __init__(self, filename):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.L = C.CDLL(REF_SO)
        self.L.ref_map_load.restype = C.c_void_p
        self.L.ref_clearance.restype = C.c_double
        self.L.ref_interpolate.restype = C.c_double
        self.L.ref_edge_nodes_batch.restype = C.c_int64
        self.L.ref_edge_index_batch.restype = C.c_int64
        self.h = C.c_void_p(self.L.ref_map_load(filename.encode()))
        if not self.h:
            raise RuntimeError("reference ESDFMap::load failed for " + filename)

    def close(self):
        if self.h:
            self.L.ref_map_free(self.h)
            self.h = None

    def min_pos(self):
        o = np.empty(3)
        self.L.ref_min_pos(self.h, _p(o))
        return o

    def max_pos(self):
        o = np.empty(3)
        self.L.ref_max_pos(self.h, _p(o))
        return o

    def clearance(self, xyz):
        xyz = _f64(xyz).reshape(-1, 3)
        out = np.empty(len(xyz))
        self.L.ref_clearance_batch(self.h, _p(xyz), C.c_int64(len(xyz)), _p(out))
        return out

    def interpolate(self, q):
        q = _f64(q)
        return self.L.ref_interpolate(self.h, _p(q))

    def gradient(self, xyz):
        xyz = _f64(xyz).reshape(-1, 3)
        g = np.empty_like(xyz)
        c = np.empty_like(xyz)
        self.L.ref_gradient_batch(self.h, _p(xyz), C.c_int64(len(xyz)), _p(g), _p(c))
        return g, c

    def in_collision(self, xyz, clr):
        xyz = _f64(xyz).reshape(-1, 3)
        out = np.empty(len(xyz), dtype=np.uint8)
        self.L.ref_in_collision_batch(self.h, _p(xyz), C.c_int64(len(xyz)), C.c_double(clr), _p(out))
        return out.astype(bool)

    def edge_nodes(self, a, b, clr, cdc, want_hit=True):
        a = _f64(a).reshape(-1, 3)
        b = _f64(b).reshape(-1, 3)
        free = np.empty(len(a), dtype=np.uint8)
        hit = np.empty((len(a), 3)) if want_hit else None
        self.L.ref_edge_nodes_batch(self.h, _p(a), _p(b), C.c_int64(len(a)), C.c_double(clr), C.c_double(cdc), _p(free), _p(hit))
        return free.astype(bool), hit

    def edge_index(self, nodes, frm, to, clr, cdc):
        nodes = _f64(nodes).reshape(-1, 3)
        frm = np.ascontiguousarray(frm, dtype=np.int32)
        to = np.ascontiguousarray(to, dtype=np.int32)
        free = np.empty(len(frm), dtype=np.uint8)
        self.L.ref_edge_index_batch(self.h, _p(nodes), _p(frm), _p(to), C.c_int64(len(frm)), C.c_double(clr),
                                    C.c_double(cdc), _p(free))
        return free.astype(bool)

    def edge_guards(self, a, b, clr, cdc):
        a = _f64(a).reshape(-1, 3)
        b = _f64(b).reshape(-1, 3)
        free = np.empty(len(a), dtype=np.uint8)
        self.L.ref_edge_guards_batch(self.h, _p(a), _p(b), C.c_int64(len(a)), C.c_double(clr), C.c_double(cdc), _p(free))
        return free.astype(bool)

    def sample_path(self, pts, length, num_samples):
        pts = _f64(pts).reshape(-1, 3)
        out = np.empty((int(num_samples), 3))
        self.L.ref_sample_path(self.h, _p(pts), C.c_int(len(pts)), C.c_double(length), C.c_double(num_samples), _p(out))
        return out

    def deformable(self, p1, len1, p2, len2, num_check, clr, cdc):
        p1 = _f64(p1).reshape(-1, 3)
        p2 = _f64(p2).reshape(-1, 3)
        return self.L.ref_deformable(self.h, _p(p1), C.c_int(len(p1)), C.c_double(len1), _p(p2), C.c_int(len(p2)),
                                     C.c_double(len2), C.c_double(num_check), C.c_double(clr), C.c_double(cdc))

    def deformable_between(self, s, e, b1, b2, clr, cdc):
        s, e, b1, b2 = (_f64(x) for x in (s, e, b1, b2))
        return self.L.ref_deformable_between(self.h, _p(s), _p(e), _p(b1), _p(b2), C.c_double(clr), C.c_double(cdc))

    def path_collision_free(self, pts, clr, cdc):
        pts = _f64(pts).reshape(-1, 3)
        hit = np.empty(3)
        r = self.L.ref_path_collision_free(self.h, _p(pts), C.c_int(len(pts)), C.c_double(clr), C.c_double(cdc), _p(hit))
        return r, hit

    def sample_multiple_batch(self, start, goal, cand, n, clr, cdc, k=13, nthreads=1, full=False):
        """ref_sample_multiple_batch: every clearance test done by the reference's own code."""
        return _sample_multiple_batch(self.L.ref_sample_multiple_batch, self.h, start, goal, cand, n, k, clr, cdc,
                                      nthreads, full, False)

    def fill_random_state_in_ellipse(self, s, g, ratio, planar, u, draws):
        s, g, draws = _f64(s), _f64(g), _f64(draws)
        self.L.ref_set_draws(C.c_double(u), _p(draws))
        o = np.empty(3)
        self.L.ref_fill_random_state_in_ellipse(self.h, _p(s), _p(g), C.c_double(ratio), C.c_int(int(planar)), _p(o))
        return o


def ref_available():
    return os.path.exists(REF_SO)
