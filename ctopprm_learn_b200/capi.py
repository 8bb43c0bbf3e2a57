"""ctypes binding of include/ctopprm_b200.h.  No compute happens here."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

LAYOUT_PLAIN, LAYOUT_CELL8, LAYOUT_TEX, LAYOUT_BRICK = 1, 2, 4, 8
LAYOUT_NAMES = {1: "plain", 2: "cell8", 4: "tex", 8: "brick"}
DEFAULT_LAYOUT = LAYOUT_TEX  # the layout bench.py uses unless told otherwise (fastest measured; n <= 1024)
EDGE_NODES, EDGE_GUARDS = 0, 1
CSR_UPPER, CSR_NO_WEIGHTS, CSR_COL16, NODES_AS_INDEX = 1, 2, 4, 8
TUNE_KNN_PER_CELL, TUNE_KNN_TILE, TUNE_PIPE_CHUNK, TUNE_EDGE_ORDER, TUNE_SSSP_FRONTIER, TUNE_PIPE_HEAD, TUNE_SSSP_CLUSTER = 0, 1, 2, 3, 4, 5, 6
TUNE_CSR_PAIR_CAP, TUNE_PERQ_CLUSTER, TUNE_SEG_BIN, TUNE_KNN_RADIUS, TUNE_EDGE_FILTER = 7, 8, 9, 10, 11
ERR_NEED_CANDIDATES = -4

_lib = None


class CtpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ctopprm_b200 error {code}: {msg}")
        self.code = code


def lib_path() -> str:
    return os.path.join(_HERE, "lib", "libctopprm_b200.so")


_P = C.c_void_p
_I64 = C.c_int64
_D = C.c_double
_I = C.c_int

# name -> argtypes (every function returns int except ctp_last_error)
_SIGS = {
    "ctp_version": [],
    "ctp_device_count": [_P],
    "ctp_map_create": [_P, _I, _P, _P, _I, _I, _P],
    "ctp_map_create_f64": [_P, _I, _P, _P, _I, _I, _P],
    "ctp_map_create_columns": [_I, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "ctp_map_create_occupancy": [_P, _I, _P, _P, _I, _I, _P, _P],
    "ctp_map_download": [_P, _P],
    "ctp_map_destroy": [_P],
    "ctp_map_set_layout": [_P, _I],
    "ctp_map_info": [_P, _P],
    "ctp_set_prm_group": [_P, _I],
    "ctp_set_tunable": [_P, _I, _D],
    "ctp_reserve": [_P, _I64, _I64, _I64, _I],
    "ctp_lookup_counter_read": [_P, _P, _P, _I],
    "ctp_filter_counter_read": [_P, _P, _P, _I],
    "ctp_launch_counter_read": [_P, _P, _I],
    "ctp_io_counter_read": [_P, _P, _P, _I],
    "ctp_profile": [_P, _I],
    "ctp_profile_read": [_P, _P, _P, _I],
    "ctp_host_alloc": [_P, C.c_size_t],
    "ctp_host_free": [_P],
    "ctp_clearance": [_P, _P, _I64, _P],
    "ctp_clearance_dev": [_P, _P, _I64, _P, _P],
    "ctp_in_collision": [_P, _P, _I64, _D, _P],
    "ctp_in_collision_dev": [_P, _P, _I64, _D, _P, _P],
    "ctp_gradient": [_P, _P, _I64, _P, _P],
    "ctp_gradient_dev": [_P, _P, _I64, _P, _P, _P],
    "ctp_edge_check": [_P, _P, _P, _I64, _D, _D, _I, _I, _P, _P, _P, _P],
    "ctp_edge_check_dev": [_P, _P, _P, _I64, _D, _D, _I, _I, _P, _P, _P, _P, _P],
    "ctp_edge_check_indexed": [_P, _P, _I64, _P, _P, _I64, _D, _D, _I, _P, _P, _P],
    "ctp_edge_check_indexed_dev": [_P, _P, _I64, _P, _P, _I64, _D, _D, _I, _P, _P, _P, _P],
    "ctp_sample_reject": [_P, _P, _P, _P, _I64, _I64, _D, _I64, _P, _P, _P, _P],
    "ctp_sample_reject_dev": [_P, _P, _P, _P, _I64, _I64, _D, _I64, _P, _P, _P, _P, _P],
    "ctp_sample_ellipsoid": [_P, _P, _P, _I64, _I64, _D, _I, C.c_uint64, _P],
    "ctp_sample_ellipsoid_dev": [_P, _P, _P, _I64, _I64, _D, _I, C.c_uint64, _P, _P],
    "ctp_knn": [_P, _P, _I64, _I64, _I, _P, _P],
    "ctp_knn_dev": [_P, _P, _I64, _I64, _I, _P, _P, _P],
    "ctp_build_prm": [_P, _P, _I64, _I64, _I, _D, _D, _P, _P, _P, _P, _P, _I64, _P],
    "ctp_build_prm_dev": [_P, _P, _I64, _I64, _I, _D, _D, _P, _P, _P, _P, _P, _I64, _P, _P],
    "ctp_build_prm_ex_dev": [_P, _P, _I64, _I64, _I, _D, _D, _I, _P, _P, _P, _P, _P, _I64, _P, _P],
    "ctp_csr_from_directed": [_P, _P, _I64, _I64, _I, _P, _P, _P, _P, _P, _I64, _P],
    "ctp_edge_check_knn_dev": [_P, _P, _I64, _I64, _I, _P, _D, _D, _P, _P],
    "ctp_prm_rows_dev": [_P, _P, _I64, _I, _D, _D, _I64, _I64, _P, _P, _P, _P],
    "ctp_csr_from_directed_dev": [_P, _P, _I64, _I64, _I, _P, _P, _P, _P, _P, _I64, _P, _P],
    "ctp_sample_multiple": [_P, _P, _P, _P, _I64, _I64, _I64, _I, _D, _D, _P, _P, _P, _P, _P, _I64, _P],
    "ctp_sample_multiple_ex": [_P, _P, _P, _P, _I64, _I64, _I64, _I, _D, _D, _I, _P, _P, _P, _P, _P, _I64, _P],
    "ctp_sample_ellipsoid_draws": [_P, _P, _P, _I64, _I64, _D, _I, C.c_uint64, _P, _P],
    "ctp_sssp": [_P, _I64, _I64, _P, _P, _P, _P, _P, _I, _P, _P, _P],
    "ctp_sssp_dev": [_P, _I64, _I64, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P],
    "ctp_cluster_flood_dev": [_P, _I64, _I64, _P, _P, _P, _P, _P, _I, _P, _I, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P],
    "ctp_roadmap_create": [_P, _I64, _I64, _P, _P, _P, _P, _I, _P],
    "ctp_roadmap_destroy": [_P],
    "ctp_roadmap_info": [_P, _P],
    "ctp_cluster_flood": [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "ctp_query_paths": [_P, _P, _P, _P, _I64, _I64, _I64, _I, _D, _D, C.c_int32, _P, _P, _P, _P, _P],
    "ctp_sample_path": [_P, _P, _P, _P, _P, _I64, _P, _P, _P],
    "ctp_deformable": [_P, _P, _P, _P, _I64, _P, _P, _P, _I64, _D, _D, _P],
    "ctp_shorten_path": [_P, _P, _P, _P, _I64, _I, _D, _D, C.c_int32, _P, _P, _P, _P, _P],
    "ctp_shorten_path_ex": [_P, _P, _P, _P, _I64, _I, _P, _D, _D, C.c_int32, _P, _P, _P, _P, _P],
}


def lib():
    """Load libctopprm_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "-- there is no CPU fallback for the CTopPRM hot path")
        L = C.CDLL(path)
        L.ctp_last_error.restype = C.c_char_p
        L.ctp_last_error.argtypes = []
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = args
        _lib = L
    return _lib


def bind_to_gpu_numa_node(device=0):
    """Pin this process (and the page-locked buffers it allocates afterwards) to the CPUs of the NUMA node the GPU hangs
    off: with one process per GPU every rank then stages its host <-> device copies through its own node's memory instead of
    all ranks sharing one.  Returns {"node", "cpus"} or None when the topology cannot be read (nothing is changed then)."""
    import torch
    try:
        pr = torch.cuda.get_device_properties(device)
        base = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(base + "/numa_node").read().strip())
        cpulist = open(base + "/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if node < 0 or not cpus or cpus == allowed:
            return {"node": node, "cpus": len(allowed), "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus), "bound": True}
    except Exception:
        return None


def pinned_empty(shape, dtype):
    """numpy array over page-locked host memory (ctp_host_alloc); freed with the array."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    p = C.c_void_p()
    _check(lib().ctp_host_alloc(C.byref(p), max(1, count * dtype.itemsize)))
    buf = (C.c_char * max(1, count * dtype.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    import weakref
    weakref.finalize(buf, lambda addr=p.value: lib().ctp_host_free(C.c_void_p(addr)))
    return arr


def nodes_from_index(node_src, cand, start, goal):
    """coordinates of the nodes a NODES_AS_INDEX call reported by position in the candidate streams
    (include/ctopprm/csr_expand.hpp: nodes_from_index)"""
    node_src = np.asarray(node_src)
    q, n = node_src.shape
    cand = np.asarray(cand, dtype=np.float64).reshape(q, -1, 3)
    nodes = np.zeros((q, n, 3))
    for i in range(q):
        s = node_src[i]
        ok = s >= 0
        nodes[i, ok] = cand[i, s[ok]]
        nodes[i, s == -2] = np.asarray(start, dtype=np.float64).reshape(-1, 3)[i]
        nodes[i, s == -3] = np.asarray(goal, dtype=np.float64).reshape(-1, 3)[i]
    return nodes


def exported_symbols():
    return ["ctp_last_error"] + list(_SIGS)


def _check(rc, allow=()):
    if rc != 0 and rc not in allow:
        raise CtpError(rc, lib().ctp_last_error().decode())
    return rc


def _np(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def _ptr(a):
    """numpy array / torch tensor / int / None -> void*"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.data_ptr())  # torch tensor


def _stream_ptr(stream):
    if stream is None:
        return None
    if isinstance(stream, int):
        return C.c_void_p(stream)
    return C.c_void_p(stream.cuda_stream)


class DeviceMap:
    """ESDF resident on one GPU.  Mirrors ESDFMap/BaseMap's batchable queries
    (src/esdf_map.cpp, src/base_map.cpp); numpy in / numpy out uses the host-pointer ABI,
    the ``*_dev`` methods take torch CUDA tensors and a stream."""

    def __init__(self, vox, centroid, extents, device=0, layouts=LAYOUT_PLAIN):
        L = lib()
        vox = np.asarray(vox)
        if vox.ndim != 3 or not (vox.shape[0] == vox.shape[1] == vox.shape[2]):
            raise ValueError("voxels must be a cubic (N,N,N) array (src/esdf_map.cpp:24)")
        self.n = int(vox.shape[0])
        self.centroid = _np(centroid, np.float64)
        self.extents = _np(extents, np.float64)
        h = C.c_void_p()
        if vox.dtype == np.float64:
            v = np.ascontiguousarray(vox)
            _check(L.ctp_map_create_f64(_ptr(v), self.n, _ptr(self.centroid), _ptr(self.extents), device, layouts, C.byref(h)))
        else:
            v = _np(vox, np.float32)
            _check(L.ctp_map_create(_ptr(v), self.n, _ptr(self.centroid), _ptr(self.extents), device, layouts, C.byref(h)))
        self._h = h
        self.device = device

    @classmethod
    def columns(cls, n, centroid, extents, cx, cy, r, device=0, layouts=LAYOUT_PLAIN):
        """procedural ESDF of vertical cylinders generated on the device (ctp_map_create_columns)"""
        self = cls.__new__(cls)
        self.n = int(n)
        self.centroid = _np(centroid, np.float64)
        self.extents = _np(extents, np.float64)
        cx, cy, r = _np(cx, np.float64), _np(cy, np.float64), _np(r, np.float64)
        h = C.c_void_p()
        _check(lib().ctp_map_create_columns(self.n, _ptr(self.centroid), _ptr(self.extents), _ptr(cx), _ptr(cy), _ptr(r),
                                            len(cx), device, layouts, C.byref(h)))
        self._h = h
        self.device = device
        return self

    @classmethod
    def from_occupancy(cls, occ, centroid, extents, device=0, layouts=LAYOUT_PLAIN, want_d2=False):
        """exact signed EDT of a voxel occupancy grid on the device (ctp_map_create_occupancy);
        with want_d2 returns (map, signed integer squared voxel distances)"""
        occ = _np(np.asarray(occ) != 0, np.uint8)
        if occ.ndim != 3 or len(set(occ.shape)) != 1:
            raise ValueError("occupancy grid must be n x n x n")
        self = cls.__new__(cls)
        self.n = int(occ.shape[0])
        self.centroid = _np(centroid, np.float64)
        self.extents = _np(extents, np.float64)
        d2 = np.empty(occ.shape, dtype=np.int32) if want_d2 else None
        h = C.c_void_p()
        _check(lib().ctp_map_create_occupancy(_ptr(occ), self.n, _ptr(self.centroid), _ptr(self.extents), device, layouts,
                                              _ptr(d2) if want_d2 else None, C.byref(h)))
        self._h = h
        self.device = device
        return (self, d2) if want_d2 else self

    def download(self):
        vox = np.empty((self.n, self.n, self.n), dtype=np.float32)
        _check(lib().ctp_map_download(self._h, _ptr(vox)))
        return vox

    @classmethod
    def load(cls, filename, device=0, layouts=LAYOUT_PLAIN):
        """ESDFMap::load (src/esdf_map.cpp:6-72): four concatenated NPY records."""
        from .synth import read_map
        if not os.path.exists(filename):
            raise RuntimeError("npy_load: Unable to open file " + filename)  # src/esdf_map.cpp:11-12
        vox, c, e, _ = read_map(filename)
        return cls(vox, c, e, device=device, layouts=layouts)

    def close(self):
        if getattr(self, "_h", None):
            lib().ctp_map_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- bookkeeping ---------------------------------------------------------------------
    def info(self):
        a = np.zeros(12)
        _check(lib().ctp_map_info(self._h, _ptr(a)))
        return a

    def layout_name(self):
        return LAYOUT_NAMES[int(self.info()[7])]

    def prm_group(self):
        return int(self.info()[10])

    def set_tunable(self, key, value):
        _check(lib().ctp_set_tunable(self._h, key, float(value)))

    def set_prm_group(self, group):
        _check(lib().ctp_set_prm_group(self._h, group))

    def get_min_pos(self):
        return self.info()[0:3]

    def get_max_pos(self):
        return self.info()[3:6]

    def set_layout(self, layout):
        _check(lib().ctp_map_set_layout(self._h, layout))

    def reserve(self, q, n, m, k=13):
        _check(lib().ctp_reserve(self._h, q, n, m, k))

    def read_lookups(self, stream=None, reset=True):
        v = C.c_uint64(0)
        _check(lib().ctp_lookup_counter_read(self._h, _stream_ptr(stream), C.byref(v), int(reset)))
        return v.value

    def read_filter_counter(self, stream=None, reset=True):
        """(edges decided by the whole-edge filter, edges it was shown, reference-order lookups of the decided ones)
        since the last reset"""
        v = (C.c_uint64 * 3)(0, 0, 0)
        _check(lib().ctp_filter_counter_read(self._h, _stream_ptr(stream), v, int(reset)))
        return int(v[0]), int(v[1]), int(v[2])

    def read_launches(self, reset=True):
        v = C.c_uint64(0)
        _check(lib().ctp_launch_counter_read(self._h, C.byref(v), int(reset)))
        return v.value

    def read_io(self, reset=True):
        """(host->device bytes, device->host bytes) moved by the chunk pipelines since the last reset"""
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(lib().ctp_io_counter_read(self._h, C.byref(a), C.byref(b), int(reset)))
        return a.value, b.value

    STAGES = ("reject", "knn", "edges", "csr")

    def profile(self, enable=True):
        _check(lib().ctp_profile(self._h, int(enable)))

    def profile_read(self, reset=True):
        """-> ({stage: device ms}, {stage: spans}) accumulated since the last reset"""
        ms = np.zeros(len(self.STAGES))
        calls = np.zeros(len(self.STAGES), dtype=np.int64)
        _check(lib().ctp_profile_read(self._h, _ptr(ms), _ptr(calls), int(reset)))
        return dict(zip(self.STAGES, ms.tolist())), dict(zip(self.STAGES, calls.tolist()))

    # -- host-pointer API ----------------------------------------------------------------
    def clearance(self, xyz):
        xyz = _np(xyz, np.float64).reshape(-1, 3)
        out = np.empty(len(xyz))
        _check(lib().ctp_clearance(self._h, _ptr(xyz), len(xyz), _ptr(out)))
        return out

    def in_collision(self, xyz, clr):
        xyz = _np(xyz, np.float64).reshape(-1, 3)
        out = np.empty(len(xyz), dtype=np.uint8)
        _check(lib().ctp_in_collision(self._h, _ptr(xyz), len(xyz), clr, _ptr(out)))
        return out.astype(bool)

    def gradient(self, xyz):
        xyz = _np(xyz, np.float64).reshape(-1, 3)
        g = np.empty_like(xyz)
        c = np.empty_like(xyz)
        _check(lib().ctp_gradient(self._h, _ptr(xyz), len(xyz), _ptr(g), _ptr(c)))
        return g, c

    def edge_check(self, a, b, clr, cdc, mode=EDGE_NODES, group=0, pinned=False, out=None):
        """out: (free u8, hit f64 x3, hit_idx i32, lookups i32) arrays to reuse (e.g. page-locked ones)"""
        a = _np(a, np.float64).reshape(-1, 3)
        b = _np(b, np.float64).reshape(-1, 3)
        m = len(a)
        if out is not None:
            free, hit, hit_idx, lookups = out
        else:
            alloc = pinned_empty if pinned else (lambda shape, dt: np.empty(shape, dtype=dt))
            free = alloc((m,), np.uint8)
            hit = alloc((m, 3), np.float64)
            hit_idx = alloc((m,), np.int32)
            lookups = alloc((m,), np.int32)
        _check(lib().ctp_edge_check(self._h, _ptr(a), _ptr(b), m, clr, cdc, mode, group, _ptr(free), _ptr(hit),
                                    _ptr(hit_idx), _ptr(lookups)))
        return free.astype(bool), hit, hit_idx, lookups

    def edge_check_indexed(self, nodes, frm, to, clr, cdc, group=0):
        nodes = _np(nodes, np.float64).reshape(-1, 3)
        frm = _np(frm, np.int32)
        to = _np(to, np.int32)
        m = len(frm)
        free = np.empty(m, dtype=np.uint8)
        hit_idx = np.empty(m, dtype=np.int32)
        lookups = np.empty(m, dtype=np.int32)
        _check(lib().ctp_edge_check_indexed(self._h, _ptr(nodes), len(nodes), _ptr(frm), _ptr(to), m, clr, cdc, group,
                                            _ptr(free), _ptr(hit_idx), _ptr(lookups)))
        return free.astype(bool), hit_idx, lookups

    def sample_reject(self, start, goal, cand, clr, n, allow_short=False):
        start = _np(start, np.float64).reshape(-1, 3)
        goal = _np(goal, np.float64).reshape(-1, 3)
        q = len(start)
        cand = _np(cand, np.float64).reshape(q, -1, 3)
        m = cand.shape[1]
        nodes = np.empty((q, n, 3))
        n_filled = np.empty(q, dtype=np.int32)
        n_used = np.empty(q, dtype=np.int32)
        accept = np.empty((q, m), dtype=np.uint8)
        _check(lib().ctp_sample_reject(self._h, _ptr(start), _ptr(goal), _ptr(cand), q, m, clr, n, _ptr(nodes),
                                       _ptr(n_filled), _ptr(n_used), _ptr(accept)),
               allow=(ERR_NEED_CANDIDATES,) if allow_short else ())
        return nodes, n_filled, n_used, accept.astype(bool)

    def sample_ellipsoid(self, start, goal, m, ratio, planar, seed):
        start = _np(start, np.float64).reshape(-1, 3)
        goal = _np(goal, np.float64).reshape(-1, 3)
        q = len(start)
        cand = np.empty((q, m, 3))
        _check(lib().ctp_sample_ellipsoid(self._h, _ptr(start), _ptr(goal), q, m, ratio, int(planar), seed, _ptr(cand)))
        return cand

    def sample_ellipsoid_draws(self, start, goal, m, ratio, planar, seed):
        """candidates plus the (u, g0, g1, g2) draws each one was made from (q x m x 4)"""
        start = _np(start, np.float64).reshape(-1, 3)
        goal = _np(goal, np.float64).reshape(-1, 3)
        q = len(start)
        cand = np.empty((q, m, 3))
        draws = np.empty((q, m, 4))
        _check(lib().ctp_sample_ellipsoid_draws(self._h, _ptr(start), _ptr(goal), q, m, ratio, int(planar), seed, _ptr(cand),
                                                _ptr(draws)))
        return cand, draws

    def knn(self, nodes, k=13):
        nodes = _np(nodes, np.float64)
        if nodes.ndim == 2:
            nodes = nodes[None]
        q, n, _ = nodes.shape
        idx = np.empty((q, n, k), dtype=np.int32)
        d2 = np.empty((q, n, k), dtype=np.float32)
        _check(lib().ctp_knn(self._h, _ptr(nodes), q, n, k, _ptr(idx), _ptr(d2)))
        return idx, d2

    def build_prm_ex(self, nodes, clr, cdc, k=13, csr_flags=0):
        """compact adjacency (CSR_UPPER / CSR_NO_WEIGHTS) through the device entry point, torch staging"""
        import torch
        nodes = _np(nodes, np.float64)
        if nodes.ndim == 2:
            nodes = nodes[None]
        q, n, _ = nodes.shape
        dev = torch.device("cuda", self.device)
        cap = 2 * q * n * k
        nd = torch.from_numpy(nodes).to(dev)
        rp = torch.zeros((q, n + 1), dtype=torch.int32, device=dev)
        col = torch.zeros(cap, dtype=torch.int32, device=dev)
        w = None if csr_flags & CSR_NO_WEIGHTS else torch.zeros(cap, dtype=torch.float64, device=dev)
        off = torch.zeros(q + 1, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream(dev)
        _check(lib().ctp_build_prm_ex_dev(self._h, _ptr(nd), q, n, k, clr, cdc, csr_flags, None, None, _ptr(rp), _ptr(col),
                                          _ptr(w), cap, _ptr(off), _stream_ptr(st)))
        torch.cuda.synchronize(dev)
        off = off.cpu().numpy()
        nnz = int(off[q])
        return dict(row_ptr=rp.cpu().numpy(), col=col[:nnz].cpu().numpy(), w=None if w is None else w[:nnz].cpu().numpy(),
                    nnz_off=off)

    def build_prm(self, nodes, clr, cdc, k=13):
        nodes = _np(nodes, np.float64)
        if nodes.ndim == 2:
            nodes = nodes[None]
        q, n, _ = nodes.shape
        cap = 2 * q * n * k
        nbr = np.empty((q, n, k), dtype=np.int32)
        dfree = np.empty((q, n, k), dtype=np.uint8)
        row_ptr = np.empty((q, n + 1), dtype=np.int32)
        col = np.empty(cap, dtype=np.int32)
        w = np.empty(cap)
        off = np.empty(q + 1, dtype=np.int64)
        _check(lib().ctp_build_prm(self._h, _ptr(nodes), q, n, k, clr, cdc, _ptr(nbr), _ptr(dfree), _ptr(row_ptr),
                                   _ptr(col), _ptr(w), cap, _ptr(off)))
        nnz = int(off[q])
        return dict(nbr=nbr, directed_free=dfree.astype(bool), row_ptr=row_ptr, col=col[:nnz], w=w[:nnz], nnz_off=off)

    def csr_from_directed(self, nodes, nbr, directed_free):
        nodes = _np(nodes, np.float64)
        if nodes.ndim == 2:
            nodes = nodes[None]
        q, n, _ = nodes.shape
        nbr = _np(nbr, np.int32).reshape(q, n, -1)
        k = nbr.shape[2]
        fr = _np(directed_free, np.uint8).reshape(q, n, k)
        cap = 2 * q * n * k
        row_ptr = np.empty((q, n + 1), dtype=np.int32)
        col = np.empty(cap, dtype=np.int32)
        w = np.empty(cap)
        off = np.empty(q + 1, dtype=np.int64)
        _check(lib().ctp_csr_from_directed(self._h, _ptr(nodes), q, n, k, _ptr(nbr), _ptr(fr), _ptr(row_ptr), _ptr(col),
                                           _ptr(w), cap, _ptr(off)))
        nnz = int(off[q])
        return dict(row_ptr=row_ptr, col=col[:nnz], w=w[:nnz], nnz_off=off)

    def sssp(self, row_ptr, col, w, nnz_off, src, want_label=False):
        """shortest paths from src (q x ns node ids) over q CSR graphs -> dist, pred[, label]"""
        row_ptr = _np(row_ptr, np.int32)
        if row_ptr.ndim == 1:
            row_ptr = row_ptr[None]
        q, n = row_ptr.shape[0], row_ptr.shape[1] - 1
        col, w = _np(col, np.int32), _np(w, np.float64)
        nnz_off = _np(nnz_off, np.int64)
        src = _np(src, np.int32).reshape(q, -1)
        dist = np.empty((q, n))
        pred = np.empty((q, n), dtype=np.int32)
        label = np.empty((q, n), dtype=np.int32) if want_label else None
        _check(lib().ctp_sssp(self._h, q, n, _ptr(row_ptr), _ptr(col), _ptr(w), _ptr(nnz_off), _ptr(src), src.shape[1],
                              _ptr(dist), _ptr(pred), _ptr(label)))
        return (dist, pred, label) if want_label else (dist, pred)

    def sssp_dev(self, q, n, row_ptr, col, w, nnz_off, src, ns, dist, pred, label=None, stream=None):
        _check(lib().ctp_sssp_dev(self._h, q, n, _ptr(row_ptr), _ptr(col), _ptr(w), _ptr(nnz_off), _ptr(src), ns,
                                  _ptr(dist), _ptr(pred), _ptr(label), _stream_ptr(stream)))

    def roadmap(self, row_ptr, col, w, nnz_off, max_seeds=16):
        """q CSR roadmaps resident on the device for the floods of the cluster stage"""
        return Roadmap(self, row_ptr, col, w, nnz_off, max_seeds)

    def sample_multiple(self, start, goal, cand, n, clr, cdc, k=13, out=None, csr_flags=0):
        """TopologicalPRMClustering::sampleMultiple from host candidates to host CSR
        (csr_flags: CSR_UPPER / CSR_NO_WEIGHTS select the compact forms of the adjacency; CSR_COL16: out["col"] is a
        uint16 array; NODES_AS_INDEX: out["nodes"] is an int32 (q, n) array of candidate positions, see
        nodes_from_index)."""
        start = _np(start, np.float64).reshape(-1, 3)
        goal = _np(goal, np.float64).reshape(-1, 3)
        q = len(start)
        cand = _np(cand, np.float64).reshape(q, -1, 3)
        m = cand.shape[1]
        cap = 2 * q * n * k
        if out is None:
            out = dict(nodes=np.empty((q, n), dtype=np.int32) if csr_flags & NODES_AS_INDEX else np.empty((q, n, 3)),
                       n_filled=np.empty(q, dtype=np.int32), row_ptr=np.empty((q, n + 1), dtype=np.int32),
                       col=np.empty(cap, dtype=np.uint16 if csr_flags & CSR_COL16 else np.int32),
                       w=None if csr_flags & CSR_NO_WEIGHTS else np.empty(cap), nnz_off=np.empty(q + 1, dtype=np.int64))
        assert out["col"].dtype == (np.uint16 if csr_flags & CSR_COL16 else np.int32)
        assert out["nodes"].dtype == (np.int32 if csr_flags & NODES_AS_INDEX else np.float64)
        cap = min(cap, out["col"].size)
        _check(lib().ctp_sample_multiple_ex(self._h, _ptr(start), _ptr(goal), _ptr(cand), q, m, n, k, clr, cdc, csr_flags,
                                            _ptr(out["nodes"]), _ptr(out["n_filled"]), _ptr(out["row_ptr"]),
                                            _ptr(out["col"]), _ptr(out.get("w")), cap, _ptr(out["nnz_off"])))
        return out

    def query_paths(self, start, goal, cand, n, clr, cdc, k=13, path_cap=512, out=None):
        """PRM build + shortest start->goal path per query; only the paths come back to the host"""
        start = _np(start, np.float64).reshape(-1, 3)
        goal = _np(goal, np.float64).reshape(-1, 3)
        q = len(start)
        cand = _np(cand, np.float64).reshape(q, -1, 3)
        m = cand.shape[1]
        if out is None:
            out = dict(path_ids=np.empty((q, path_cap), dtype=np.int32), path_xyz=np.empty((q, path_cap, 3)),
                       path_n=np.empty(q, dtype=np.int32), length=np.empty(q), n_filled=np.empty(q, dtype=np.int32))
        _check(lib().ctp_query_paths(self._h, _ptr(start), _ptr(goal), _ptr(cand), q, m, n, k, clr, cdc, path_cap,
                                     _ptr(out["path_ids"]), _ptr(out["path_xyz"]), _ptr(out["path_n"]),
                                     _ptr(out["length"]), _ptr(out["n_filled"])))
        return out

    @staticmethod
    def _pack_polylines(paths):
        off = np.zeros(len(paths) + 1, dtype=np.int32)
        for i, p in enumerate(paths):
            off[i + 1] = off[i] + len(p)
        pts = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 3) for p in paths]))
        return pts, off

    def sample_path(self, paths, lengths, num_samples):
        pts, off = self._pack_polylines(paths)
        lengths = _np(lengths, np.float64)
        ns = _np(num_samples, np.float64)
        cnt = len(paths)
        oo = np.zeros(cnt + 1, dtype=np.int64)
        oo[1:] = np.cumsum(ns.astype(np.int64))
        out = np.empty((int(oo[-1]), 3))
        st = np.empty(cnt, dtype=np.int32)
        _check(lib().ctp_sample_path(self._h, _ptr(pts), _ptr(off), _ptr(lengths), _ptr(ns), cnt, _ptr(out), _ptr(oo), _ptr(st)))
        return [out[oo[i]:oo[i + 1]] for i in range(cnt)], st

    def deformable(self, paths, lengths, ia, ib, num_check, clr, cdc):
        pts, off = self._pack_polylines(paths)
        lengths = _np(lengths, np.float64)
        ia = _np(ia, np.int32)
        ib = _np(ib, np.int32)
        nc = _np(num_check, np.float64)
        out = np.empty(len(ia), dtype=np.uint8)
        _check(lib().ctp_deformable(self._h, _ptr(pts), _ptr(off), _ptr(lengths), len(paths), _ptr(ia), _ptr(ib),
                                    _ptr(nc), len(ia), clr, cdc, _ptr(out)))
        return out

    def shorten_path(self, paths, lengths, forward, clr, cdc, cap=None):
        """forward: one flag for the batch, or a sequence with one flag per polyline (ctp_shorten_path_ex)"""
        pts, off = self._pack_polylines(paths)
        lengths = _np(lengths, np.float64)
        cnt = len(paths)
        if cap is None:
            cap = int(max(len(p) for p in paths) + 64)
        out = np.empty((cnt, cap, 3))
        origin = np.empty((cnt, cap), dtype=np.int32)
        out_n = np.empty(cnt, dtype=np.int32)
        out_len = np.empty(cnt)
        st = np.empty(cnt, dtype=np.int32)
        if np.ndim(forward) == 0:
            _check(lib().ctp_shorten_path(self._h, _ptr(pts), _ptr(off), _ptr(lengths), cnt, int(forward), clr, cdc, cap,
                                          _ptr(out), _ptr(origin), _ptr(out_n), _ptr(out_len), _ptr(st)))
        else:
            each = _np(np.asarray(forward).astype(np.uint8), np.uint8)
            assert len(each) == cnt
            _check(lib().ctp_shorten_path_ex(self._h, _ptr(pts), _ptr(off), _ptr(lengths), cnt, 0, _ptr(each), clr, cdc, cap,
                                             _ptr(out), _ptr(origin), _ptr(out_n), _ptr(out_len), _ptr(st)))
        return [out[i, :out_n[i]] for i in range(cnt)], [origin[i, :out_n[i]] for i in range(cnt)], out_len, st

    # -- device-pointer API (torch CUDA tensors) -----------------------------------------
    def clearance_dev(self, xyz, out, stream=None):
        _check(lib().ctp_clearance_dev(self._h, _ptr(xyz), xyz.numel() // 3, _ptr(out), _stream_ptr(stream)))

    def edge_check_dev(self, a, b, clr, cdc, free, mode=EDGE_NODES, group=0, hit=None, hit_idx=None, lookups=None,
                       stream=None):
        _check(lib().ctp_edge_check_dev(self._h, _ptr(a), _ptr(b), a.numel() // 3, clr, cdc, mode, group, _ptr(free),
                                        _ptr(hit), _ptr(hit_idx), _ptr(lookups), _stream_ptr(stream)))

    def edge_check_indexed_dev(self, nodes, n_nodes, frm, to, clr, cdc, free, group=0, hit_idx=None, lookups=None, stream=None):
        _check(lib().ctp_edge_check_indexed_dev(self._h, _ptr(nodes), n_nodes, _ptr(frm), _ptr(to), frm.numel(), clr, cdc,
                                                group, _ptr(free), _ptr(hit_idx), _ptr(lookups), _stream_ptr(stream)))

    def sample_reject_dev(self, start, goal, cand, q, m, clr, n, nodes, n_filled, n_used, accept=None, stream=None):
        _check(lib().ctp_sample_reject_dev(self._h, _ptr(start), _ptr(goal), _ptr(cand), q, m, clr, n, _ptr(nodes),
                                           _ptr(n_filled), _ptr(n_used), _ptr(accept), _stream_ptr(stream)))

    def sample_ellipsoid_dev(self, start, goal, q, m, ratio, planar, seed, cand, stream=None):
        _check(lib().ctp_sample_ellipsoid_dev(self._h, _ptr(start), _ptr(goal), q, m, ratio, int(planar), seed,
                                              _ptr(cand), _stream_ptr(stream)))

    def knn_dev(self, nodes, q, n, k, idx, d2=None, stream=None):
        _check(lib().ctp_knn_dev(self._h, _ptr(nodes), q, n, k, _ptr(idx), _ptr(d2), _stream_ptr(stream)))

    def csr_from_directed_dev(self, nodes, q, n, k, nbr, directed_free, row_ptr, col, w, cap, nnz_off, stream=None):
        _check(lib().ctp_csr_from_directed_dev(self._h, _ptr(nodes), q, n, k, _ptr(nbr), _ptr(directed_free), _ptr(row_ptr),
                                               _ptr(col), _ptr(w), cap, _ptr(nnz_off), _stream_ptr(stream)))

    def prm_rows_dev(self, nodes, n, k, clr, cdc, first, last, nbr, directed_free, stream=None):
        """neighbour search + directed checks for the slice [first, last) of the cell order -> rows done"""
        done = C.c_int64(0)
        _check(lib().ctp_prm_rows_dev(self._h, _ptr(nodes), n, k, clr, cdc, first, last, _ptr(nbr), _ptr(directed_free),
                                      C.byref(done), _stream_ptr(stream)))
        return done.value

    def edge_check_knn_dev(self, nodes, q, n, k, nbr, clr, cdc, directed_free, stream=None):
        _check(lib().ctp_edge_check_knn_dev(self._h, _ptr(nodes), q, n, k, _ptr(nbr), clr, cdc, _ptr(directed_free),
                                            _stream_ptr(stream)))

    def build_prm_dev(self, nodes, q, n, k, clr, cdc, row_ptr, col, w, cap, nnz_off, nbr=None, directed_free=None,
                      stream=None, csr_flags=0):
        _check(lib().ctp_build_prm_ex_dev(self._h, _ptr(nodes), q, n, k, clr, cdc, csr_flags, _ptr(nbr), _ptr(directed_free),
                                          _ptr(row_ptr), _ptr(col), _ptr(w), cap, _ptr(nnz_off), _stream_ptr(stream)))


class Roadmap:
    """ctp_roadmap: q roadmaps on the device plus the state of their cluster floods (dist / prev / label)."""

    def __init__(self, dmap, row_ptr, col, w, nnz_off, max_seeds=16):
        row_ptr = _np(row_ptr, np.int32)
        if row_ptr.ndim == 1:
            row_ptr = row_ptr[None]
        self.q, self.n = row_ptr.shape[0], row_ptr.shape[1] - 1
        self.max_seeds = int(max_seeds)
        self._map = dmap  # keeps the map handle alive
        col, w, nnz_off = _np(col, np.int32), _np(w, np.float64), _np(nnz_off, np.int64)
        h = C.c_void_p()
        _check(lib().ctp_roadmap_create(dmap._h, self.q, self.n, _ptr(row_ptr), _ptr(col), _ptr(w), _ptr(nnz_off),
                                        self.max_seeds, C.byref(h)))
        self._h = h
        info = np.zeros(4, dtype=np.int64)
        _check(lib().ctp_roadmap_info(self._h, _ptr(info)))
        self.rec_cap = int(info[3])

    def flood(self, seeds, mode=0):
        """one flood (mode 0: all seeds, mode 1: the newest seed on top of the previous state); seeds is a list of
        per-query seed lists.  -> dist, prev, label (q x n), records [(m_i x 2 nodes, m_i dist, m_i labels)], n_updates (q)"""
        sd = np.full((self.q, self.max_seeds), -1, dtype=np.int32)
        ns = np.zeros(self.q, dtype=np.int32)
        for i, lst in enumerate(seeds):
            ns[i] = len(lst)
            sd[i, :len(lst)] = lst
        dist = np.empty((self.q, self.n))
        prev = np.empty((self.q, self.n), dtype=np.int32)
        label = np.empty((self.q, self.n), dtype=np.int32)
        rn = np.empty((self.q, self.rec_cap, 2), dtype=np.int32)
        rd = np.empty((self.q, self.rec_cap))
        rl = np.empty((self.q, self.rec_cap), dtype=np.int32)
        rc = np.zeros(self.q, dtype=np.int32)
        upd = np.zeros(self.q, dtype=np.uint64)
        _check(lib().ctp_cluster_flood(self._h, _ptr(sd), _ptr(ns), mode, _ptr(dist), _ptr(prev), _ptr(label), _ptr(rn),
                                       _ptr(rl), _ptr(rd), _ptr(rc), _ptr(upd)))
        recs = [(rn[i, :rc[i]].copy(), rd[i, :rc[i]].copy(), rl[i, :rc[i]].copy()) for i in range(self.q)]
        return dist, prev, label, recs, upd

    def close(self):
        if getattr(self, "_h", None):
            lib().ctp_roadmap_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
