"""ctypes view of libctopprm_host.so: the C++ class surface of include/ctopprm/*.hpp (ESDFMap, BaseMap,
TopologicalPRMClustering) driven through the flat test entry points of csrc/host/host_capi.cpp.
Used by the parity tests; a C++ caller includes the headers directly."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

DEFAULT_YAML = """\
# builder-authored defaults (BASELINE.md section 4); the reference ships no YAML
map_type: ESDF
map: {map}
min_clearance: {clr}
check_collisions: true
collision_distance_check: {cdc}
ellipse_ratio_major_axis_focal_length: 1.2
planar: {planar}
num_samples_between_gate: {n}
num_clusters: 2
min_clusters: 2
max_clusters: 10
min_ratio: 1.1
max_path_length_ratio: 1.5
cutoof_distance_ratio_to_shortest: 1.3
start:
  position: [{s[0]}, {s[1]}, {s[2]}]
end:
  position: [{g[0]}, {g[1]}, {g[2]}]
"""


def lib_path():
    return os.path.join(_HERE, "lib", "libctopprm_host.so")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path()):
            raise ImportError(lib_path() + " is missing: run __graft_entry__.build()")
        L = C.CDLL(lib_path())
        L.ctph_map_load.restype = C.c_void_p
        L.ctph_map_from_occupancy.restype = C.c_void_p
        L.ctph_map_from_memory.restype = C.c_void_p
        L.ctph_planner_new.restype = C.c_void_p
        L.ctph_get_clearence.restype = C.c_double
        L.ctph_planner_graph.restype = C.c_int64
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def yaml_get(text, key):
    buf = C.create_string_buffer(256)
    ok = lib().ctph_yaml_get(text.encode(), key.encode(), buf, 256)
    return buf.value.decode() if ok else None


class HostMap:
    """ESDFMap (include/ctopprm/esdf_map.hpp) loaded from a map FILE."""

    def __init__(self, filename, device=0, layouts=0):
        self.h = C.c_void_p(lib().ctph_map_load(filename.encode(), device, layouts))
        if not self.h:
            raise RuntimeError("ESDFMap::load failed for " + filename)

    @classmethod
    def from_occupancy(cls, occ, centroid, extents, device=0, layouts=0):
        """ESDFMap::loadFromOccupancy: exact signed distance transform of a voxel occupancy grid on the device"""
        occ = np.ascontiguousarray(np.asarray(occ) != 0, dtype=np.uint8)
        c, e = _f(centroid), _f(extents)
        self = cls.__new__(cls)
        self.h = C.c_void_p(lib().ctph_map_from_occupancy(_p(occ), occ.shape[0], _p(c), _p(e), device, layouts))
        return self

    @classmethod
    def from_memory(cls, vox, centroid, extents, device=0, layouts=0):
        """ESDFMap::loadFromMemory: the voxel grid as the .npy record holds it (x*n*n + y*n + z)"""
        vox = np.ascontiguousarray(vox, dtype=np.float32)
        c, e = _f(centroid), _f(extents)
        self = cls.__new__(cls)
        self.h = C.c_void_p(lib().ctph_map_from_memory(_p(vox), vox.shape[0], _p(c), _p(e), device, layouts))
        return self

    def save(self, filename):
        """ESDFMap::save: the four NPY records of map.py:124-132"""
        if lib().ctph_map_save(self.h, filename.encode()) != 0:
            raise RuntimeError("ESDFMap::save failed for " + filename)

    def close(self):
        if self.h:
            lib().ctph_map_free(self.h)
            self.h = None

    def get_clearence(self, p):
        p = _f(p)
        return lib().ctph_get_clearence(self.h, _p(p))

    def min_max(self):
        o = np.empty(6)
        lib().ctph_min_max(self.h, _p(o))
        return o[:3], o[3:]

    def gradient(self, p):
        p = _f(p)
        g, c = np.empty(3), np.empty(3)
        lib().ctph_gradient(self.h, _p(p), _p(g), _p(c))
        return g, c

    def real_pos_from_index(self, q):
        q = _f(q)
        o = np.empty(3)
        lib().ctph_real_pos_from_index(self.h, _p(q), _p(o))
        return o

    def in_collision(self, p, clr):
        p = _f(p)
        return bool(lib().ctph_in_collision(self.h, _p(p), C.c_double(clr)))

    def edge_nodes(self, a, b, clr, cdc):
        a, b = _f(a), _f(b)
        hit = np.empty(3)
        r = lib().ctph_edge_nodes(self.h, _p(a), _p(b), C.c_double(clr), C.c_double(cdc), _p(hit))
        return bool(r), hit

    def edge_guards(self, a, b, clr, cdc):
        a, b = _f(a), _f(b)
        return bool(lib().ctph_edge_guards(self.h, _p(a), _p(b), C.c_double(clr), C.c_double(cdc)))

    def path_collision_free(self, pts, clr, cdc):
        pts = _f(pts).reshape(-1, 3)
        hit = np.empty(3)
        r = lib().ctph_path_collision_free(self.h, _p(pts), len(pts), C.c_double(clr), C.c_double(cdc), _p(hit))
        return bool(r), hit

    def deformable(self, p1, l1, p2, l2, num, clr, cdc):
        p1, p2 = _f(p1).reshape(-1, 3), _f(p2).reshape(-1, 3)
        return lib().ctph_deformable(self.h, _p(p1), len(p1), C.c_double(l1), _p(p2), len(p2), C.c_double(l2),
                                     C.c_double(num), C.c_double(clr), C.c_double(cdc))

    def deformable_between(self, s, e, b1, b2, clr, cdc):
        s, e, b1, b2 = (_f(x) for x in (s, e, b1, b2))
        return lib().ctph_deformable_between(self.h, _p(s), _p(e), _p(b1), _p(b2), C.c_double(clr), C.c_double(cdc))

    def sample_path(self, pts, length, num):
        pts = _f(pts).reshape(-1, 3)
        out = np.empty((int(num), 3))
        lib().ctph_sample_path(self.h, _p(pts), len(pts), C.c_double(length), C.c_double(num), _p(out))
        return out

    def fill_random_state_in_ellipse(self, s, g, ratio, planar, seed, count):
        s, g = _f(s), _f(g)
        out = np.empty((count, 3))
        lib().ctph_fill_random_state_in_ellipse(self.h, _p(s), _p(g), C.c_double(ratio), int(planar), C.c_uint64(seed),
                                                count, _p(out))
        return out

    def find_geometrical_paths(self, yaml_text, start, goal, out_dir, cap=64, cand=None, cap_pts=1 << 16, want_paths=False):
        """TopologicalPRMClustering::find_geometrical_paths for one start-goal pair (out_dir "" = no CSV files);
        cand = the candidate stream of sampleMultiple (else drawn on the device).  -> lengths[, paths]"""
        s, g = _f(start), _f(goal)
        lens = np.empty(cap)
        pts = np.empty((cap_pts, 3))
        off = np.zeros(cap + 1, dtype=np.int32)
        c = None if cand is None else _f(cand).reshape(-1, 3)
        n = lib().ctph_find_geometrical_paths(self.h, yaml_text.encode(), _p(s), _p(g), out_dir.encode(),
                                              None if c is None else _p(c), C.c_int64(0 if c is None else len(c)), cap,
                                              cap_pts, _p(pts), _p(off), _p(lens))
        if n < 0:
            raise RuntimeError("find_geometrical_paths: output capacity too small")
        if want_paths:
            return lens[:n].copy(), [pts[off[i]:off[i + 1]].copy() for i in range(n)]
        return lens[:n].copy()


class HostPlanner:
    """TopologicalPRMClustering<Vector<3>> (include/ctopprm/topological_prm_clustering.hpp)."""

    def __init__(self, hmap, yaml_text, start, goal):
        s, g = _f(start), _f(goal)
        self.map = hmap
        self.h = C.c_void_p(lib().ctph_planner_new(hmap.h, yaml_text.encode(), _p(s), _p(g)))
        if not self.h:
            raise RuntimeError("planner construction failed")

    def close(self):
        if self.h:
            lib().ctph_planner_free(self.h)
            self.h = None

    def set_candidates(self, cand):
        cand = _f(cand).reshape(-1, 3)
        lib().ctph_planner_set_candidates(self.h, _p(cand), C.c_int64(len(cand)))

    def sample_multiple(self, n):
        lib().ctph_planner_sample_multiple(self.h, n)
        self.n = n

    def graph(self, k=13):
        n = self.n
        nodes = np.empty((n, 3))
        row_ptr = np.empty(n + 1, dtype=np.int32)
        cap = 2 * n * k + 64
        col = np.empty(cap, dtype=np.int32)
        w = np.empty(cap)
        nnz = lib().ctph_planner_graph(self.h, _p(nodes), _p(row_ptr), _p(col), _p(w), C.c_int64(cap))
        assert nnz >= 0
        return nodes, row_ptr, col[:nnz], w[:nnz]

    def save_roadmaps(self, dense_file, csr_file):
        lib().ctph_planner_save_roadmaps(self.h, dense_file.encode(), csr_file.encode())

    def flood(self, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.int32)
        dist = np.empty(self.n)
        pred = np.empty(self.n, dtype=np.int32)
        cluster = np.empty(self.n, dtype=np.int32)
        lib().ctph_planner_flood(self.h, _p(seeds), len(seeds), _p(dist), _p(pred), _p(cluster))
        return dist, pred, cluster

    def cluster_graph(self, max_path_length, seed_candidates=(), cap_paths=4096, cap_pts=1 << 20):
        """clusterGraph (:1832-1952) -> (paths, DFS lengths, seeds, labels, stats)"""
        sc = np.ascontiguousarray(seed_candidates, dtype=np.int32)
        pts = np.empty((cap_pts, 3))
        off = np.zeros(cap_paths + 1, dtype=np.int32)
        lens = np.empty(cap_paths)
        seeds = np.full(1024, -1, dtype=np.int32)
        ns = C.c_int32(0)
        labels = np.empty(self.n, dtype=np.int32)
        stats = np.zeros(5, dtype=np.int32)
        k = lib().ctph_planner_cluster_graph(self.h, C.c_double(max_path_length), _p(sc), len(sc), cap_paths, cap_pts, _p(pts),
                                             _p(off), _p(lens), _p(seeds), C.byref(ns), _p(labels), _p(stats))
        if k < 0:
            raise RuntimeError("cluster_graph: output capacity too small")
        names = ("seeds", "floods", "refloods", "capped_refloods", "homotopy_checks")
        return ([pts[off[i]:off[i + 1]].copy() for i in range(k)], lens[:k].copy(), seeds[:ns.value].copy(), labels,
                dict(zip(names, (int(v) for v in stats))))

    def seed_accepted(self, candidate_id, seed_ids):
        seed_ids = np.ascontiguousarray(seed_ids, dtype=np.int32)
        return bool(lib().ctph_planner_seed_accepted(self.h, int(candidate_id), _p(seed_ids), len(seed_ids)))

    def shortest(self, cap=4096):
        ids = np.empty(cap, dtype=np.int32)
        length = C.c_double(0)
        cnt = lib().ctph_planner_shortest(self.h, _p(ids), cap, C.byref(length))
        return ids[:min(cnt, cap)].copy(), length.value


def last_phase_times():
    """phases of the last find_geometrical_paths call (ms) and the cluster stage's counters"""
    t = np.zeros(7)
    c = np.zeros(5, dtype=np.int32)
    lib().ctph_last_phase_times(_p(t), _p(c))
    names = ("prm", "dijkstra", "wavefront", "min_tours", "dfs_and_stitch", "shorten_and_filter", "total")
    out = {k: float(v) for k, v in zip(names, t)}
    out.update({k: int(v) for k, v in zip(("cluster_paths", "seeds", "refloods", "homotopy_checks", "capped_refloods"), c)})
    return out


def shorten_path(pts, length, forward, cap=512):
    pts = _f(pts).reshape(-1, 3)
    out = np.empty((cap, 3))
    out_len = C.c_double(0)
    n = lib().ctph_shorten_path(_p(pts), len(pts), C.c_double(length), int(forward), _p(out), cap, C.byref(out_len))
    return out[:n].copy(), out_len.value


def remove_equivalent(paths, lengths):
    off = np.zeros(len(paths) + 1, dtype=np.int32)
    for i, p in enumerate(paths):
        off[i + 1] = off[i] + len(p)
    pts = _f(np.concatenate([np.asarray(p).reshape(-1, 3) for p in paths]))
    lengths = _f(lengths)
    keep = np.empty(len(paths), dtype=np.int32)
    n = lib().ctph_remove_equivalent(_p(pts), _p(off), _p(lengths), len(paths), _p(keep))
    return keep[:n].copy()


def new_homotopy_batch(min_paths, max_paths, min_lens, max_lens):
    paths = list(min_paths) + list(max_paths)
    off = np.zeros(len(paths) + 1, dtype=np.int32)
    for i, p in enumerate(paths):
        off[i + 1] = off[i] + len(p)
    pts = _f(np.concatenate([np.asarray(p).reshape(-1, 3) for p in paths]))
    lens = _f(list(min_lens) + list(max_lens))
    out = np.empty(len(min_paths), dtype=np.uint8)
    lib().ctph_new_homotopy_batch(_p(pts), _p(off), _p(lens), len(min_paths), _p(out))
    return out


def remove_too_long(lengths):
    lengths = _f(lengths)
    keep = np.empty(len(lengths), dtype=np.int32)
    n = lib().ctph_remove_too_long(_p(lengths), len(lengths), _p(keep))
    return keep[:n].copy()


def csr_expand_upper(nodes, row_ptr_up, col_up, off_up, threads=1, want_w=True, out=None):
    """host-side mirror (include/ctopprm/csr_expand.hpp): compact adjacency -> full symmetric weighted CSR.
    nodes q x n x 3, row_ptr_up q x (n+1), col_up packed, off_up q+1.  out: dict(row_ptr, col, w, nnz_off) to reuse."""
    nodes = _f(nodes)
    q, n = nodes.shape[0], nodes.shape[1]
    rp = np.ascontiguousarray(row_ptr_up, dtype=np.int32)
    cu = np.ascontiguousarray(col_up, dtype=np.int32)
    ou = np.ascontiguousarray(off_up, dtype=np.int64)
    nnz = 2 * int(ou[q])
    if out is None:
        out = dict(row_ptr=np.empty((q, n + 1), dtype=np.int32), col=np.empty(nnz, dtype=np.int32),
                   w=np.empty(nnz) if want_w else None, nnz_off=np.empty(q + 1, dtype=np.int64))
    w = out.get("w")
    lib().ctph_csr_expand_upper(_p(nodes), C.c_longlong(q), C.c_longlong(n), _p(rp), _p(cu), _p(ou), _p(out["row_ptr"]),
                                _p(out["col"]), _p(w) if w is not None else None, _p(out["nnz_off"]), C.c_int(threads))
    return out
