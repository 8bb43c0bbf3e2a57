#!/usr/bin/env python
"""bench.py -- edge collision checks/s of the CTopPRM PRM-build hot path on B200.

One step = one pass of TopologicalPRMClustering::sampleMultiple
(include/topological_prm_clustering.hpp:288-401: rejection sampling -> k-NN(13) -> n*13 directed
segment marches -> symmetric CSR) over a batch of Q independent start-goal queries on the C2 map
(random columns, 256^3 voxels @ 0.05 m, 10 000 nodes per query -- BASELINE.json configs[1]; the
batch is the per-GPU shard of configs[3]).  See DESIGN.md section "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                          # the reference's CPU code
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ctopprm_learn_b200 import synth  # noqa: E402

METRIC = "edge collision checks/s"
UNIT = "edge checks/s"
K_NN = 13
WORKLOADS = {
    # name: (map config, nodes per query, default queries per GPU per step, candidates per node)
    "C2": ("C2", 10000, 512, 4),
    "C1": ("C1", 1000, 512, 4),
    "C3": ("C3", 100000, 4, 16),
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def make_queries(c, e, q_total, n, cand_per_node, lo, hi, out=None, cfg=None):
    """C4 query set (seed 4): start / goal ~U in free space of the map `cfg`; rank slice [lo, hi).
    Candidate streams: seed 400000 + query id."""
    s, g = synth.random_queries(c, e, q_total, 4, cfg=cfg)
    s, g = s[lo:hi], g[lo:hi]
    m = cand_per_node * n
    cand = out if out is not None else np.empty((hi - lo, m, 3))
    for i in range(hi - lo):
        cand[i] = synth.ellipsoid_candidates(s[i], g[i], m, 400000 + lo + i)
    return np.ascontiguousarray(s), np.ascontiguousarray(g), cand


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx = float(f[2])
                except ValueError:
                    continue
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=mx, samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# The contract is ONE JSON line on stdout.  Libraries below us write there too (NCCL prints its version banner on
# stdout when NCCL_DEBUG asks for it), so file descriptor 1 is pointed at stderr for the whole run and the line
# goes out through a private duplicate of the original stdout.
_REAL_STDOUT = None


def guard_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


EPSILON_BAND = 1e-9  # north_star: verdicts may differ only where |clearance - min_clearance| < epsilon at the deciding sample


def detailed_parity(dmap, om, nodes_d, n, pq, clr, cdc, dev, torch):
    """Array-by-array comparison of the first pq roadmaps with the CPU restatement (oracle/), counted on the device:
    k-NN lists, per-directed-edge verdicts, first-hit march indices, CSR columns and FP64 weights.  A verdict mismatch
    is 'explained' when the clearance of some march sample of that edge lies within EPSILON_BAND of the threshold."""
    k = K_NN
    cap = 2 * pq * n * k
    nbr_d = torch.zeros((pq, n, k), dtype=torch.int32, device=dev)
    fr_d = torch.zeros((pq, n, k), dtype=torch.uint8, device=dev)
    rp_d = torch.zeros((pq, n + 1), dtype=torch.int32, device=dev)
    col_d = torch.zeros(cap, dtype=torch.int32, device=dev)
    w_d = torch.zeros(cap, dtype=torch.float64, device=dev)
    off_d = torch.zeros(pq + 1, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream()
    dmap.build_prm_dev(nodes_d[:pq], pq, n, k, clr, cdc, rp_d, col_d, w_d, cap, off_d, nbr=nbr_d, directed_free=fr_d, stream=st)
    rows = torch.arange(pq * n, dtype=torch.int32, device=dev).repeat_interleave(k)
    base = (torch.arange(pq, dtype=torch.int32, device=dev) * n).repeat_interleave(n * k)
    nb = nbr_d.reshape(-1)
    to = torch.where(nb >= 0, nb + base, nb)
    free2 = torch.zeros(pq * n * k, dtype=torch.uint8, device=dev)
    hidx_d = torch.zeros(pq * n * k, dtype=torch.int32, device=dev)
    dmap.edge_check_indexed_dev(nodes_d[:pq], pq * n, rows, to, clr, cdc, free2, hit_idx=hidx_d, stream=st)
    torch.cuda.synchronize()
    nodes_h = nodes_d[:pq].cpu().numpy()
    want = [om.build_prm(nodes_h[i], clr, cdc, k=k, nthreads=cpu_cores()) for i in range(pq)]
    frm_h = np.repeat(np.arange(n, dtype=np.int32), k)
    hidx_h = np.stack([om.edge_index(nodes_h[i], frm_h, want[i]["nbr"].reshape(-1), clr, cdc, nthreads=cpu_cores())[1]
                       for i in range(pq)])
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    nbr_w = T(np.stack([w_["nbr"] for w_ in want]))
    fr_w = T(np.stack([w_["directed_free"] for w_ in want]).astype(np.uint8))
    knn_mis = int((nbr_w != nbr_d).sum().item())
    bad = (fr_w != fr_d)
    directed_mis = int(bad.sum().item())
    first_hit_mis = int((T(hidx_h).reshape(-1) != hidx_d).sum().item())
    rp_w = T(np.stack([w_["row_ptr"] for w_ in want]))
    rowptr_mis = int((rp_w != rp_d).sum().item())
    col_mis = w_mis = 0
    off_h = off_d.cpu().numpy()
    for i in range(pq):
        a, b = int(off_h[i]), int(off_h[i + 1])
        if b - a != len(want[i]["col"]):
            col_mis += abs(b - a - len(want[i]["col"]))
            continue
        col_mis += int((T(want[i]["col"]) != col_d[a:b]).sum().item())
        w_mis += int((T(want[i]["w"]) != w_d[a:b]).sum().item())
    in_band = 0
    if directed_mis:
        idx = torch.nonzero(bad.reshape(-1)).reshape(-1).cpu().numpy()
        nbr_h = nbr_d.cpu().numpy().reshape(pq, n, k)
        for e in idx:
            qi, r, sl = int(e // (n * k)), int((e // k) % n), int(e % k)
            a_, b_ = nodes_h[qi][r], nodes_h[qi][nbr_h[qi, r, sl]]
            npts = np.ceil(np.sqrt(((b_ - a_) ** 2).sum()) / cdc + 1.0)
            t = np.arange(1, int(npts)) / npts
            cl = om.clearance(a_ + (b_ - a_) * t[:, None])
            in_band += int(np.nanmin(np.abs(cl - clr)) < EPSILON_BAND) if len(cl) else 0
    return {"queries_compared_array_by_array": int(pq), "directed_edges_compared": int(pq * n * k),
            "knn_index_mismatch": knn_mis, "directed_mismatch": directed_mis, "epsilon": EPSILON_BAND,
            "epsilon_band": in_band, "directed_mismatch_outside_epsilon_band": directed_mis - in_band,
            "first_hit_mismatch": first_hit_mis, "row_ptr_mismatch": rowptr_mis, "csr_col_mismatch": col_mis,
            "csr_weight_mismatch": w_mis,
            "against": "oracle/ctp_oracle.c restatement (exact float32 k-NN in place of FLANN's approximate search), "
                       "compared on the device"}


class CpuArm:
    """The reference's CPU implementation of the path: oracle/_ref (the reference's own
    src/esdf_map.cpp + src/base_map.cpp) when it was built, else the oracle port."""

    def __init__(self, vox, c, e):
        from oracle import orc
        self.orc = orc
        orc.build()
        self.kind = "reference" if orc.ref_available() else "port"
        self.om = orc.OracleMap(vox, c, e)
        self.ref = None
        self.tmp = None
        if self.kind == "reference":
            fd, self.tmp = tempfile.mkstemp(suffix=".npy")
            os.close(fd)
            synth.write_map(self.tmp, vox, c, e)
            self.ref = orc.RefMap(self.tmp)  # ESDFMap::load, nested-vector grid as include/esdf_map.hpp:43

    def run(self, s, g, cand, n, clr, cdc, threads):
        t0 = time.perf_counter()
        if self.ref is not None:
            out = self.ref.sample_multiple_batch(s, g, cand, n, clr, cdc, k=K_NN, nthreads=threads)
        else:
            out = self.orc.sample_multiple_batch(self.om, s, g, cand, n, clr, cdc, k=K_NN, nthreads=threads)
        return time.perf_counter() - t0, out

    def close(self):
        if self.ref is not None:
            self.ref.close()
        if self.tmp:
            os.unlink(self.tmp)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on the host cores, rank 0 only."""
    if rank != 0:
        return
    mapcfg, n, qdef, cpn = WORKLOADS[args.workload]
    vox, c, e = synth.make_config_map(mapcfg)
    clr, cdc = synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK
    cores = cpu_cores()
    arm = CpuArm(vox, c, e)
    q_total = (args.queries or qdef) * world
    # bounded sample: one query per core per step (at least 8), the first queries of the same set
    q_s = args.cpu_queries or max(8, min(4 * cores, 512))
    s, g, cand = make_queries(c, e, q_total, n, cpn, 0, min(q_s, q_total), cfg=mapcfg)
    q_s = len(s)
    for _ in range(args.warmup):
        arm.run(s[:max(1, q_s // 8)], g[:max(1, q_s // 8)], cand[:max(1, q_s // 8)], n, clr, cdc, cores)
    t = 0.0
    for _ in range(args.steps):
        dt, out = arm.run(s, g, cand, n, clr, cdc, cores)
        t += dt
    checks = q_s * n * K_NN * args.steps
    val = checks / t
    sample = f"{q_s} of {q_total} queries per step (n={n}, {n * K_NN} directed checks each), OpenMP over queries"
    line = {
        "metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, n, args.queries or qdef, cpn, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ms_per_query": 1e3 * t / args.steps / q_s * min(cores, q_s),
        "free_edge_fraction": float(out["n_free"].sum() / (q_s * n * K_NN)),
    }
    arm.close()
    emit_line(line)


def workload_config(name, n, q, cpn, world):
    mapcfg = WORKLOADS[name][0]
    kind, extent, nvox, nobs, seed, _ = synth.CONFIGS[mapcfg]
    return {
        "workload": f"{name}: {kind} ESDF {nvox}^3 voxels @ {extent / nvox:.3f} m (seed {seed}), {n}-node PRM per query "
                    f"(k-NN {K_NN} -> {n * K_NN} directed edge checks), {q} independent start-goal queries per step per GPU "
                    f"(C4 query set, seed 4), sampleMultiple = reject + kNN + edge march + CSR",
        "queries_per_step_per_gpu": q, "nodes_per_query": n, "k": K_NN, "candidates_per_query": cpn * n,
        "min_clearance": synth.MIN_CLEARANCE, "collision_distance_check": synth.COLLISION_DISTANCE_CHECK,
        "parallelism": f"queries sharded over {world} GPU(s), map replicated, no data-path collective",
        "l2": "per-step inputs (candidate streams) exceed the 126 MB L2; the ESDF grid is re-read within a step by design",
    }


def run_paths_leg(dmap, nodes_d, row_ptr_d, col_d, w_d, nnz_off_d, nq, n, clr, cdc):
    """Shortest roadmap paths of the first nq queries (host Dijkstra over the CSR handed back, as the
    reference does), then one ctp_shorten_path batch per direction and one all-pairs ctp_deformable batch
    per query group -- the batched forms of shorten_path / removeEquivalentPaths."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra
    nodes = nodes_d[:nq].cpu().numpy()
    rp = row_ptr_d[:nq].cpu().numpy()
    off = nnz_off_d[:nq + 1].cpu().numpy()
    lo, hi = int(off[0]), int(off[nq])
    col = col_d[lo:hi].cpu().numpy()
    w = w_d[lo:hi].cpu().numpy()
    paths, lens = [], []
    for i in range(nq):
        a, b = int(off[i]) - lo, int(off[i + 1]) - lo
        g = csr_matrix((w[a:b], col[a:b], rp[i]), shape=(n, n))
        dist, pred = dijkstra(g, directed=True, indices=0, return_predecessors=True)
        if not np.isfinite(dist[1]):
            continue
        p = [1]
        while p[-1] != 0:
            p.append(int(pred[p[-1]]))
        pts = nodes[i][p[::-1]]
        paths.append(pts)
        lens.append(float(np.sum(np.sqrt(((pts[1:] - pts[:-1]) ** 2).sum(1)))))
    if len(paths) < 2:
        return None
    cap = max(len(p) for p in paths) + 64
    dmap.shorten_path(paths, lens, False, clr, cdc, cap=cap)  # warm-up
    t0 = time.perf_counter()
    back = dmap.shorten_path(paths, lens, False, clr, cdc, cap=cap)
    fwd = dmap.shorten_path(back[0], list(back[2]), True, clr, cdc, cap=cap)
    t_short = time.perf_counter() - t0
    t0 = time.perf_counter()  # second pass: best of two (a wall-clock millisecond is easily doubled by a host hiccup)
    back = dmap.shorten_path(paths, lens, False, clr, cdc, cap=cap)
    fwd = dmap.shorten_path(back[0], list(back[2]), True, clr, cdc, cap=cap)
    t_short = min(t_short, time.perf_counter() - t0)
    sp, sl = fwd[0], list(fwd[2])
    P = len(sp)
    ia = [i for i in range(P) for j in range(i + 1, P)]
    ib = [j for i in range(P) for j in range(i + 1, P)]
    nc = [float(np.ceil(max(sl[i], sl[j]) / cdc) + 1) for i, j in zip(ia, ib)]
    dmap.deformable(sp, sl, ia, ib, nc, clr, cdc)  # warm-up (same size: the scratch arena grows on first use)
    t0 = time.perf_counter()
    d = dmap.deformable(sp, sl, ia, ib, nc, clr, cdc)
    t_def = time.perf_counter() - t0
    t0 = time.perf_counter()
    d = dmap.deformable(sp, sl, ia, ib, nc, clr, cdc)
    t_def = min(t_def, time.perf_counter() - t0)
    return {"polylines": P, "mean_nodes_per_path": float(np.mean([len(p) for p in paths])),
            "shorten_both_directions_ms": 1e3 * t_short, "shortened_mean_nodes": float(np.mean([len(p) for p in sp])),
            "length_ratio_after_shortening": float(np.sum(sl) / np.sum(lens)),
            "deformable_pairs": len(ia), "deformable_ms": 1e3 * t_def, "deformable_true": int((d == 1).sum()),
            "connectors_checked_upper_bound": int(np.sum(nc)),
            "api": "ctp_shorten_path / ctp_deformable (host buffers, wall clock)"}


def run_legs(args):
    """The other BASELINE configs as legs of the default line (N = 1): each one is this script run as a child process
    on the same GPU with its own workload flag; the child's JSON line is condensed here."""
    import subprocess as sp
    legs = {}
    todo = {
        "C3": ["--workload", "C3", "--steps", "20", "--no-paths", "--no-sustained", "--no-e2e-full"],
        "MU_C2": ["--workload", "MU", "--mu-map", "C2", "--steps", "10"],
        "MU_C3": ["--workload", "MU", "--mu-map", "C3", "--steps", "10"],
        "C5": ["--workload", "C5", "--steps", "10"],
    }
    keep = ("value", "unit", "ms_per_step", "steps", "roofline", "parity", "cpu_baseline", "e2e", "stage_ms_per_step",
            "free_edge_fraction", "lookups_per_s", "shortest_path", "ms_per_query", "gpu_launches")
    for name, flags in todo.items():
        t0 = time.perf_counter()
        try:
            r = sp.run([sys.executable, os.path.abspath(__file__), "--gpus", "1", "--warmup", "3", "--no-legs"] + flags,
                       stdout=sp.PIPE, stderr=sp.PIPE, timeout=900, env=dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK=os.environ.get("LOCAL_RANK", "0")))
            out = r.stdout.decode().strip().splitlines()
            d = json.loads(out[-1]) if out else {}
            leg = {k_: d[k_] for k_ in keep if k_ in d}
            leg["workload"] = d.get("config", {}).get("workload")
            if r.returncode != 0 or not d:
                leg["error"] = r.stderr.decode()[-400:]
        except Exception as ex:  # a leg must never take the headline down with it
            leg = {"error": repr(ex)[:400]}
        leg["wall_s"] = time.perf_counter() - t0
        legs[name] = leg
    return legs


def run_multimap(args, rank, world):
    """--workload C5 (BASELINE configs[4]): a sweep over procedurally generated maps, one handle and one batch of
    queries per map.  64 maps per GPU per step (seeds 100 + 64*rank ...), random columns 256^3 as C2, generated
    on the device; 8 queries of 10 000 nodes per map.  Same metric: directed edge checks / s over the whole
    sampleMultiple."""
    import torch
    import torch.distributed as dist
    from ctopprm_learn_b200 import capi

    local = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    kind, extent, nvox, n_col, _, n = synth.CONFIGS["C2"]
    n_maps, qpm, cpn = args.maps or 64, 8, 4
    clr, cdc = synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK
    e = np.array([extent] * 3)
    m_c = cpn * n
    maps, s_l, g_l, cand_l = [], [], [], []
    t0 = time.perf_counter()
    for mi in range(n_maps):
        seed = 100 + rank * n_maps + mi
        c, cx, cy, r = synth.column_params(n_col, extent, seed)
        dm = capi.DeviceMap.columns(nvox, c, e, cx, cy, r, device=local, layouts=capi.LAYOUT_PLAIN | capi.LAYOUT_TEX)
        dm.set_layout(capi.LAYOUT_TEX)
        dm.reserve(qpm, n, m_c, K_NN)
        maps.append(dm)
        s, g = synth.random_queries(c, e, qpm, seed, cfg=((cx, cy, r), None))
        s_l.append(s)
        g_l.append(g)
        cand_l.append(np.stack([synth.ellipsoid_candidates(s[i], g[i], m_c, seed * 1000 + i) for i in range(qpm)]))
    t_setup = time.perf_counter() - t0
    q = n_maps * qpm
    T = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    s_d, g_d = T(np.concatenate(s_l), None), T(np.concatenate(g_l), None)
    cand_d = T(np.concatenate(cand_l), None)
    cap = 2 * qpm * n * K_NN
    nodes_d = torch.zeros((q, n, 3), dtype=torch.float64, device=dev)
    nf_d = torch.zeros(q, dtype=torch.int32, device=dev)
    nu_d = torch.zeros(q, dtype=torch.int32, device=dev)
    rp_d = torch.zeros((q, n + 1), dtype=torch.int32, device=dev)
    col_d = torch.zeros((n_maps, cap), dtype=torch.int32, device=dev)
    w_d = torch.zeros((n_maps, cap), dtype=torch.float64, device=dev)
    off_d = torch.zeros((n_maps, qpm + 1), dtype=torch.int64, device=dev)
    streams = [torch.cuda.Stream() for _ in range(max(1, args.map_streams))]
    main = torch.cuda.current_stream()

    nbr_d = torch.zeros((q, n, K_NN), dtype=torch.int32, device=dev)
    fr_d = torch.zeros((q, n, K_NN), dtype=torch.uint8, device=dev)
    col_all = torch.zeros(n_maps * cap, dtype=torch.int32, device=dev)
    w_all = torch.zeros(n_maps * cap, dtype=torch.float64, device=dev)
    off_all = torch.zeros(q + 1, dtype=torch.int64, device=dev)
    maps[0].reserve(q, n, m_c, K_NN)  # the handle that runs the map-independent stages for the whole step

    def fan_out():
        ev = torch.cuda.Event()
        ev.record(main)
        for st in streams:
            st.wait_event(ev)

    def fan_in():
        for st in streams:
            e2 = torch.cuda.Event()
            e2.record(st)
            main.wait_event(e2)

    def step_per_map():
        fan_out()
        for mi, dm in enumerate(maps):
            st = streams[mi % len(streams)]
            a, b = mi * qpm, (mi + 1) * qpm
            dm.sample_reject_dev(s_d[a:b], g_d[a:b], cand_d[a:b], qpm, m_c, clr, n, nodes_d[a:b], nf_d[a:b], nu_d[a:b], stream=st)
            dm.build_prm_dev(nodes_d[a:b], qpm, n, K_NN, clr, cdc, rp_d[a:b], col_d[mi], w_d[mi], cap, off_d[mi], stream=st)
        fan_in()

    def step_staged():
        # rejection and the directed checks map by map; the neighbour and adjacency stages never look at a map and
        # run once over all queries of the step (ctp_knn_dev / ctp_edge_check_knn_dev / ctp_csr_from_directed_dev)
        fan_out()
        for mi, dm in enumerate(maps):
            st = streams[mi % len(streams)]
            a, b = mi * qpm, (mi + 1) * qpm
            dm.sample_reject_dev(s_d[a:b], g_d[a:b], cand_d[a:b], qpm, m_c, clr, n, nodes_d[a:b], nf_d[a:b], nu_d[a:b], stream=st)
        fan_in()
        maps[0].knn_dev(nodes_d, q, n, K_NN, nbr_d, stream=main)
        fan_out()
        for mi, dm in enumerate(maps):
            st = streams[mi % len(streams)]
            a, b = mi * qpm, (mi + 1) * qpm
            dm.edge_check_knn_dev(nodes_d[a:b], qpm, n, K_NN, nbr_d[a:b], clr, cdc, fr_d[a:b], stream=st)
        fan_in()
        maps[0].csr_from_directed_dev(nodes_d, q, n, K_NN, nbr_d, fr_d, rp_d, col_all, w_all, n_maps * cap, off_all, stream=main)

    step = step_per_map if args.c5_per_map else step_staged

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    if int(nf_d.min()) < n:
        raise SystemExit("bench.py: a query ran out of candidates")
    for dm in maps:
        dm.read_lookups()
        dm.read_launches()
    # optional: capture each map's batch once and replay the 64 graphs over the same streams (18 launches of 8
    # queries each per map; measured 12.0 vs 12.1 ms per step -- the sweep is bound by the small batches, not by launches)
    launches_per_step = None
    if args.map_graphs:
        step()
        sync_all()
        launches_per_step = sum(dm.read_launches() for dm in maps)
        graphs = []
        cap_stream = torch.cuda.Stream()
        for mi, dm in enumerate(maps):
            a, b = mi * qpm, (mi + 1) * qpm
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=cap_stream):
                cs = torch.cuda.current_stream()
                dm.sample_reject_dev(s_d[a:b], g_d[a:b], cand_d[a:b], qpm, m_c, clr, n, nodes_d[a:b], nf_d[a:b], nu_d[a:b], stream=cs)
                dm.build_prm_dev(nodes_d[a:b], qpm, n, K_NN, clr, cdc, rp_d[a:b], col_d[mi], w_d[mi], cap, off_d[mi], stream=cs)
            graphs.append(gr)
        eager_step = step

        def step():  # noqa: F811
            ev = torch.cuda.Event()
            ev.record(main)
            for st in streams:
                st.wait_event(ev)
            for mi, gr in enumerate(graphs):
                with torch.cuda.stream(streams[mi % len(streams)]):
                    gr.replay()
            for st in streams:
                e2 = torch.cuda.Event()
                e2.record(st)
                main.wait_event(e2)
        rp_ref = rp_d.clone()
        rp_d.zero_()
        for _ in range(3):
            step()
        sync_all()
        if not torch.equal(rp_ref, rp_d):
            raise SystemExit("bench.py: graph replay produced different roadmaps")
        for dm in maps:
            dm.read_lookups()
            dm.read_launches()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record(main)
    for _ in range(args.steps):
        step()
    ev1.record(main)
    sync_all()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    lookups = sum(dm.read_lookups() for dm in maps)
    launches = sum(dm.read_launches() for dm in maps)
    if launches_per_step is not None:
        launches = launches_per_step * args.steps  # replayed graph nodes: the same kernels, counted when captured
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    checks = q * n * K_NN
    value = checks * world * args.steps / (ms_max * 1e-3)
    # end to end: per map, host candidate streams -> host CSR
    e2e = None
    if not args.no_e2e:
        from concurrent.futures import ThreadPoolExecutor
        n_thr = max(1, args.map_threads)
        outs = [dict(nodes=capi.pinned_empty((qpm, n, 3), np.float64), n_filled=capi.pinned_empty((qpm,), np.int32),
                     row_ptr=capi.pinned_empty((qpm, n + 1), np.int32), col=capi.pinned_empty((cap,), np.int32),
                     w=capi.pinned_empty((cap,), np.float64), nnz_off=capi.pinned_empty((qpm + 1,), np.int64))
                for _ in range(n_thr)]
        pins = []
        for mi in range(n_maps):
            pc = capi.pinned_empty(cand_l[mi].shape, np.float64)
            pc[:] = cand_l[mi]
            pins.append(pc)
        pool = ThreadPoolExecutor(n_thr)

        def e2e_worker(t):
            # distinct handles may be driven from distinct host threads; ctypes drops the GIL
            for mi in range(t, n_maps, n_thr):
                maps[mi].sample_multiple(s_l[mi], g_l[mi], pins[mi], n, clr, cdc, k=K_NN, out=outs[t])

        def e2e_step():
            list(pool.map(e2e_worker, range(n_thr)))
        e2e_step()
        sync_all()
        for dm in maps:
            dm.read_io()
        k2 = 2
        t0 = time.perf_counter()
        for _ in range(k2):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        io = [dm.read_io() for dm in maps]
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": checks * world * k2 / dt, "unit": UNIT, "h2d_bytes_per_step": int(sum(x[0] for x in io) // k2),
               "d2h_bytes_per_step": int(sum(x[1] for x in io) // k2), "ms_per_step": 1e3 * dt / k2, "steps": k2,
               "api": f"ctp_sample_multiple per map (host candidate streams -> host CSR), {n_thr} host threads over the "
                      "handles, wall clock, max over ranks"}
        pool.shutdown()
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = min(cpu_cores(), qpm)  # the reference arm threads over the queries of one map
        tc, bad, done = 0.0, 0, 0
        for mi, dm in enumerate(maps):  # every map of the step, until ~20 s of host work
            arm = CpuArm(dm.download(), synth.column_params(n_col, extent, 100 + mi)[0], e)
            t1, oc = arm.run(s_l[mi], g_l[mi], cand_l[mi], n, clr, cdc, cores)
            arm.close()
            tc += t1
            done += 1
            gpu_nnz = rp_d[mi * qpm:(mi + 1) * qpm, -1].cpu().numpy().astype(np.int64)
            bad += int((gpu_nnz != oc["nnz"]).sum())
            if tc > 20.0:
                break
        cpu = {"value": done * qpm * n * K_NN / tc, "unit": UNIT, "cores": cores, "kind": arm.kind,
               "sample": f"the first {done} maps of the step (grids downloaded from the device), {qpm} queries each, whole "
                         f"sampleMultiple on the host, OpenMP over queries, {tc:.1f} s"}
        parity = {"queries_checked": done * qpm, "nnz_mismatch": bad}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"C5: multi-map sweep, {n_maps} procedurally generated random-columns ESDFs per GPU per step "
                                       f"({nvox}^3 @ {extent / nvox:.3f} m, seeds 100.., generated on the device), {qpm} queries of {n} "
                                       f"nodes per map, one device handle per map, {len(streams)} streams"
                                       + (", each map's batch replayed from a CUDA graph" if args.map_graphs else "")
                                       + (", whole build per map" if (args.c5_per_map or args.map_graphs) else
                                          "; rejection and directed checks map by map, the map-independent neighbour and "
                                          "adjacency stages once over all queries of the step"),
                           "maps_per_step_per_gpu": n_maps, "queries_per_map": qpm, "nodes_per_query": n, "k": K_NN,
                           "parallelism": f"maps sharded over {world} GPU(s), no data-path collective",
                           "l2": f"{n_maps} grids of 64 MiB are swept per step: {n_maps * 64} MiB >> 126 MB L2",
                           "map_setup_s": t_setup},
                "lookups_per_s": lookups / (ms_max * 1e-3), "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
                "cpu_baseline": cpu, "parity": parity}
        emit_line(line)
    for dm in maps:
        dm.close()
    if world > 1:
        dist.destroy_process_group()


def run_microbench(args, rank, world):
    """--workload MU (SURVEY 8d / BASELINE.md section 4, row "mu"): the gather kernel alone.  2^24 random directed
    segments per GPU per step, uniform start, isotropic direction, length U[0.2, 1.5] m, on the C2 map (--mu-map C3
    for the 4 GiB grid), one ctp_edge_check_dev call per step.  Same metric: directed edge checks / s."""
    import torch
    import torch.distributed as dist
    from ctopprm_learn_b200 import capi

    local = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    clr, cdc = synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK
    vox, c, e = synth.make_config_map(args.mu_map)
    layouts = {"auto": capi.DEFAULT_LAYOUT, "plain": capi.LAYOUT_PLAIN, "cell8": capi.LAYOUT_CELL8, "tex": capi.LAYOUT_TEX,
               "brick": capi.LAYOUT_BRICK}
    layout = layouts[args.layout]
    extra = capi.LAYOUT_CELL8 if args.layout == "auto" and vox.nbytes > (100 << 20) and not args.no_cell8 else 0
    dmap = capi.DeviceMap(vox, c, e, device=local, layouts=layout | capi.LAYOUT_PLAIN | extra)
    dmap.set_layout(layout)
    if args.seg_bin != -1:
        dmap.set_tunable(capi.TUNE_SEG_BIN, args.seg_bin)
    if args.edge_filter >= 0:
        dmap.set_tunable(capi.TUNE_EDGE_FILTER, args.edge_filter)
    m_e = args.mu_edges
    a_h, b_h = synth.random_edges(c, e, m_e, 1000 + rank)
    a_d, b_d = torch.from_numpy(a_h).to(dev), torch.from_numpy(b_h).to(dev)
    free_d = torch.zeros(m_e, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        dmap.edge_check_dev(a_d, b_d, clr, cdc, free_d, group=1, stream=stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dmap.read_lookups()
    dmap.read_launches()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    lookups = dmap.read_lookups()
    launches = dmap.read_launches()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = m_e * world * args.steps / (ms_max * 1e-3)
    free_frac = float(free_d.float().mean().item())
    # end to end: host segments in, host verdicts out (pinned)
    e2e = None
    if not args.no_e2e:
        pa, pb = capi.pinned_empty(a_h.shape, np.float64), capi.pinned_empty(b_h.shape, np.float64)
        pa[:], pb[:] = a_h, b_h
        outs = (capi.pinned_empty((m_e,), np.uint8), capi.pinned_empty((m_e, 3), np.float64),
                capi.pinned_empty((m_e,), np.int32), capi.pinned_empty((m_e,), np.int32))
        dmap.edge_check(pa, pb, clr, cdc, out=outs)
        t0 = time.perf_counter()
        for _ in range(3):
            fr = dmap.edge_check(pa, pb, clr, cdc, out=outs)[0]
        dt = (time.perf_counter() - t0) / 3
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        # bytes from the arrays the call copies: two endpoints in; verdict, first-hit point, its index, lookup count out
        e2e = {"value": m_e * world / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": int(m_e * 48),
               "d2h_bytes_per_step": int(m_e * (1 + 24 + 4 + 4)), "ms_per_step": 1e3 * float(tt.item()),
               "api": "ctp_edge_check (host segments -> host verdicts, first-hit points and indices), wall clock"}
        if not np.array_equal(fr, free_d.cpu().numpy().astype(bool)):
            raise SystemExit("bench.py: host and device entry points disagree")
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from concurrent.futures import ThreadPoolExecutor
        cores = cpu_cores()
        arm = CpuArm(vox, c, e)
        m_s = min(m_e, 1 << 21)
        bounds = np.linspace(0, m_s, cores + 1).astype(np.int64)

        def part(i):
            lo, hi = int(bounds[i]), int(bounds[i + 1])
            if arm.ref is not None:
                return arm.ref.edge_nodes(a_h[lo:hi], b_h[lo:hi], clr, cdc, want_hit=False)[0]
            return arm.om.edge_nodes(a_h[lo:hi], b_h[lo:hi], clr, cdc, nthreads=1, full=False)[0]
        t0 = time.perf_counter()
        one = part(0)
        t1 = time.perf_counter() - t0
        with ThreadPoolExecutor(cores) as pool:
            t0 = time.perf_counter()
            parts = list(pool.map(part, range(cores)))
            tc = time.perf_counter() - t0
        want = np.concatenate(parts)
        got = free_d[:m_s].cpu().numpy().astype(bool)
        cpu = {"value": m_s / tc, "unit": UNIT, "cores": cores, "kind": arm.kind,
               "sample": f"first {m_s} segments of the step, isSimplePathFreeBetweenNodes one by one, {cores} host threads "
                         f"over slices, {tc:.1f} s", "single_thread_value": len(one) / t1}
        parity = {"segments_checked": int(m_s), "verdict_mismatch": int((want != got).sum())}
        arm.close()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6553.0))
        ms_launch = ms_max / args.steps
        achieved = lookups / args.steps * 32.0 / (ms_launch * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_launch, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"MU: gather-kernel microbench, {m_e} random directed segments per GPU per step (length "
                                       f"U[0.2, 1.5] m) on the {args.mu_map} map ({vox.shape[0]}^3 voxels), one ctp_edge_check_dev "
                                       "call per step", "segments_per_step_per_gpu": m_e,
                           "layout": dmap.layout_name() + (" set; the library reads the CELL8 records for large unordered batches on grids beyond the L2" if extra and args.seg_bin == -1 else ""),
                           "seg_bin": args.seg_bin,
                           "min_clearance": clr, "collision_distance_check": cdc,
                           "l2": f"endpoints ({m_e * 48 >> 20} MiB per step) exceed the 126 MB L2; the grid is re-read by design"},
                "lookups_per_s": lookups / (ms_max * 1e-3), "lookups_per_segment": lookups / args.steps / m_e,
                "free_edge_fraction": free_frac,
                "roofline": {"bound": "hbm", "kernel": "k_edge_flat", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs",
                             "ms_per_launch": ms_launch},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "cpu_baseline": cpu, "parity": parity}
        emit_line(line)
    dmap.close()
    if world > 1:
        dist.destroy_process_group()


def run_shard_query(args, rank, world):
    """--shard-query (second regime): ONE query of the workload's size split over all GPUs --
    candidate blocks -> accept counts -> NCCL all-gather of the accepted node sets -> each GPU's slice of the neighbour
    search and of the directed checks -> MAX all-reduce of the tables -> CSR on every GPU (ctopprm_learn_b200/sharded.py,
    device tensors end to end).  Strong scaling: the work is fixed, the GPUs share it.  A step = one whole query;
    per-phase times come from a second, instrumented set of steps (a device synchronisation after every phase)."""
    import torch
    import torch.distributed as dist
    from ctopprm_learn_b200 import capi, sharded

    local = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, **({} if "RANK" in os.environ else
                                                      {"rank": 0, "world_size": 1, "init_method": "tcp://127.0.0.1:29655"}))
    mapcfg, n_def, _, cpn = WORKLOADS[args.workload]
    n = args.nodes or n_def
    clr, cdc = synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK
    vox, c, e = synth.make_config_map(mapcfg)
    layout = capi.DEFAULT_LAYOUT if vox.shape[0] <= 1024 else capi.LAYOUT_PLAIN
    extra = capi.LAYOUT_CELL8 if vox.nbytes > (100 << 20) and not args.no_cell8 else 0
    dmap = capi.DeviceMap(vox, c, e, device=local, layouts=layout | capi.LAYOUT_PLAIN | extra)
    dmap.set_layout(layout)
    be = sharded.DeviceBackend(dmap)
    s, g = synth.default_query(c, e, cfg=mapcfg)
    cand = synth.ellipsoid_candidates(s, g, cpn * n, 4242)
    lo, hi = sharded.shard_range(len(cand), rank, world)
    block = torch.from_numpy(cand[lo:hi]).to(dev)  # this rank's block of the shared sample file, resident in HBM

    def step(tm=None):
        return sharded.sample_multiple_sharded(be, s, g, block, n, clr, cdc, k=K_NN, local_device=local, timings=tm, to_host=False)

    for _ in range(max(args.warmup, 3)):
        out = step()
    torch.cuda.synchronize()
    dist.barrier()
    dmap.read_launches()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    clk = clocks.stop() if rank == 0 else None
    launches = dmap.read_launches()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    tm = {}
    for _ in range(args.steps):
        step(tm)
    phases = torch.tensor([tm[k_] for k_ in ("reject", "allgather_nodes", "knn_and_edges", "allreduce_tables", "csr")],
                          dtype=torch.float64, device=dev)
    dist.all_reduce(phases, op=dist.ReduceOp.MAX)
    phases = (phases / args.steps * 1e3).tolist()
    # identical on every rank, and identical to the one-call build on one GPU
    nnz = int(out["nnz_off"][1])
    sig = torch.stack([out["col"][:nnz].to(torch.int64).sum(), out["row_ptr"].to(torch.int64).sum(),
                       out["directed_free"].to(torch.int64).sum(), torch.tensor(nnz, device=dev)])
    sigs = torch.empty(world * 4, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sigs, sig)
    same_on_all = bool((sigs.view(world, 4) == sigs.view(world, 4)[0]).all())
    equal_one_gpu = None
    if rank == 0:
        cap = 2 * n * K_NN
        rp = torch.zeros((1, n + 1), dtype=torch.int32, device=dev)
        col = torch.zeros(cap, dtype=torch.int32, device=dev)
        w = torch.zeros(cap, dtype=torch.float64, device=dev)
        off = torch.zeros(2, dtype=torch.int64, device=dev)
        dmap.build_prm_dev(out["nodes"].reshape(1, n, 3).contiguous(), 1, n, K_NN, clr, cdc, rp, col, w, cap, off,
                           stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        equal_one_gpu = bool(int(off[1]) == nnz and torch.equal(rp[0], out["row_ptr"]) and torch.equal(col[:nnz], out["col"][:nnz])
                             and torch.equal(w[:nnz], out["w"][:nnz]))
        ms_q = ms_max / args.steps
        line = {"metric": METRIC, "value": n * K_NN / (ms_q * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_q, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{args.workload} --shard-query: ONE {n}-node query on the {mapcfg} map ({vox.shape[0]}^3 voxels) "
                                       f"split over {world} GPU(s): candidate blocks, all-gather of node sets, row slices of k-NN + "
                                       f"{n * K_NN} directed checks, MAX all-reduce of the tables, CSR replicated",
                           "nodes_per_query": n, "k": K_NN, "candidates": len(cand), "parallelism": f"one query over {world} GPU(s), NCCL"},
                "ms_per_query": ms_q,
                "phase_ms_max_over_ranks": dict(zip(("reject", "allgather_nodes", "knn_and_edges", "allreduce_tables", "csr"), phases)),
                "phase_note": "instrumented steps (device synchronised after every phase); allgather_nodes includes the host read "
                              "of the accept counts, allreduce_tables moves n*k*5 bytes",
                "identical_on_all_ranks": same_on_all, "identical_to_one_gpu_build": equal_one_gpu,
                "gpu_launches": int(launches), "clocks": clk}
        emit_line(line)
    dmap.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS) + ["C5", "MU"])
    ap.add_argument("--mu-map", default="C2", choices=["C1", "C2", "C3"], help="MU: map of the gather-kernel microbench")
    ap.add_argument("--mu-edges", type=int, default=1 << 24, help="MU: random segments per GPU per step")
    ap.add_argument("--maps", type=int, default=0, help="C5: maps per GPU per step (default 64)")
    ap.add_argument("--map-streams", type=int, default=4, help="C5: streams the per-map batches are spread over")
    ap.add_argument("--c5-per-map", action="store_true", help="C5: whole build per map (ctp_build_prm_dev) instead of the staged sweep")
    ap.add_argument("--map-graphs", action="store_true", help="C5: replay each map's batch from a CUDA graph (measured: no gain, the sweep is not launch-bound)")
    ap.add_argument("--map-threads", type=int, default=4, help="C5 e2e: host threads driving the map handles")
    ap.add_argument("--queries", type=int, default=0, help="queries per GPU per step (default: per workload)")
    ap.add_argument("--no-numa", action="store_true", help="N > 1: do not bind each rank to its GPU's NUMA node")
    ap.add_argument("--seg-bin", type=int, default=-1, help="MU: bin the segments by voxel block first (1), never (0), library default (-1)")
    ap.add_argument("--shard-query", action="store_true", help="one query of the workload's size split over all GPUs (strong scaling)")
    ap.add_argument("--nodes", type=int, default=0, help="--shard-query: nodes of the query (default: the workload's)")
    ap.add_argument("--layout", default="auto", choices=["auto", "plain", "cell8", "tex", "brick"])
    ap.add_argument("--group", type=int, default=0, help="lanes per edge in the march kernel (0 = library default)")
    ap.add_argument("--knn-per-cell", type=float, default=0.0)
    ap.add_argument("--knn-ring", action="store_true", help="ring-search k-NN only (no tile kernel)")
    ap.add_argument("--knn-mode", type=int, default=-1, help="0 per-thread ring search, 1 direct search (default)")
    ap.add_argument("--edge-filter", type=int, default=-1, help="flat edge kernel: 1 float32 verdict filter in front of the FP64 code (default), 0 FP64 only")
    ap.add_argument("--edge-order", type=int, default=-1, help="1 cell-sorted row order in the edge kernel (default), 0 node order")
    ap.add_argument("--pipe-chunk", type=int, default=0, help="queries per chunk of the e2e copy/compute pipeline")
    ap.add_argument("--cpu-queries", type=int, default=0, help="size of the CPU baseline sample (queries)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-full", action="store_true", help="skip the full-format companion of the e2e leg")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained-clock leg")
    ap.add_argument("--no-legs", action="store_true", help="skip the C3 / MU / C5 legs of the default line")
    ap.add_argument("--parity-queries", type=int, default=16, help="queries compared array by array with the oracle")
    ap.add_argument("--no-single", action="store_true", help="skip the single-query latency leg")
    ap.add_argument("--no-paths", action="store_true", help="skip the shortening / mutual-visibility leg")
    ap.add_argument("--no-sssp", action="store_true", help="skip the shortest-path leg")
    ap.add_argument("--no-cell8", action="store_true", help="large grids: do not materialise the CELL8 records for the rejection stage")
    ap.add_argument("--pipe-head", type=float, default=0.0, help="candidates uploaded up front, as a multiple of n")
    ap.add_argument("--no-planner", action="store_true", help="skip the C++ class-surface single-query leg")
    ap.add_argument("--e2e-steps", type=int, default=0)
    args = ap.parse_args()
    guard_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.workload == "MU" and args.impl != "reference":
        run_microbench(args, rank, world)
        return
    if args.shard_query and args.impl != "reference":
        if args.workload not in WORKLOADS:
            raise SystemExit("bench.py: --shard-query needs --workload C1, C2 or C3")
        run_shard_query(args, rank, world)
        return
    if args.workload == "MU":
        args.workload = "C2"  # no separate CPU arm for the microbench: its own cpu_baseline leg carries the comparison
    if args.workload == "C5":
        if args.impl == "reference":
            args.workload = "C2"  # the CPU arm of the sweep is the C2 arm: same grid size, nodes and checks per query
        else:
            run_multimap(args, rank, world)
            return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from ctopprm_learn_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CTopPRM hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # one process per GPU: host staging buffers and the threads that fill them stay on the GPU's own NUMA node
    numa = capi.bind_to_gpu_numa_node(local) if world > 1 and not args.no_numa else None

    mapcfg, n, qdef, cpn = WORKLOADS[args.workload]
    q = args.queries or qdef
    clr, cdc = synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK
    vox, c, e = synth.make_config_map(mapcfg)
    layouts = {"auto": capi.LAYOUT_PLAIN, "plain": capi.LAYOUT_PLAIN, "cell8": capi.LAYOUT_CELL8,
               "tex": capi.LAYOUT_TEX, "brick": capi.LAYOUT_BRICK}
    layout = layouts[args.layout]
    if args.layout == "auto":
        layout = capi.DEFAULT_LAYOUT if vox.shape[0] <= 1024 else capi.LAYOUT_PLAIN
    extra = 0
    if args.layout == "auto" and vox.nbytes > (100 << 20) and not args.no_cell8:
        extra = capi.LAYOUT_CELL8  # grid beyond the L2: one-sector records for the scattered lookups of the rejection
    dmap = capi.DeviceMap(vox, c, e, device=local, layouts=layout | capi.LAYOUT_PLAIN | extra)
    dmap.set_layout(layout)
    if args.group:
        dmap.set_prm_group(args.group)
    if args.knn_per_cell:
        dmap.set_tunable(capi.TUNE_KNN_PER_CELL, args.knn_per_cell)
    if args.edge_order >= 0:
        dmap.set_tunable(capi.TUNE_EDGE_ORDER, args.edge_order)
    if args.edge_filter >= 0:
        dmap.set_tunable(capi.TUNE_EDGE_FILTER, args.edge_filter)
    if args.pipe_chunk:
        dmap.set_tunable(capi.TUNE_PIPE_CHUNK, args.pipe_chunk)
    if args.pipe_head:
        dmap.set_tunable(capi.TUNE_PIPE_HEAD, args.pipe_head)
    if args.knn_mode >= 0:
        dmap.set_tunable(capi.TUNE_KNN_TILE, args.knn_mode)
    if args.knn_ring:
        dmap.set_tunable(capi.TUNE_KNN_TILE, 0)

    # this rank's shard of the query set, staged in page-locked host memory
    m_c = cpn * n
    cand_h = capi.pinned_empty((q, m_c, 3), np.float64)
    s_h, g_h, _ = make_queries(c, e, q * world, n, cpn, rank * q, (rank + 1) * q, out=cand_h, cfg=mapcfg)

    stream = torch.cuda.current_stream()
    cand_d = torch.from_numpy(cand_h).to(dev)
    s_d, g_d = torch.from_numpy(s_h).to(dev), torch.from_numpy(g_h).to(dev)
    cap = 2 * q * n * K_NN
    nodes_d = torch.zeros((q, n, 3), dtype=torch.float64, device=dev)
    n_filled_d = torch.zeros(q, dtype=torch.int32, device=dev)
    n_used_d = torch.zeros(q, dtype=torch.int32, device=dev)
    row_ptr_d = torch.zeros((q, n + 1), dtype=torch.int32, device=dev)
    col_d = torch.zeros(cap, dtype=torch.int32, device=dev)
    w_d = torch.zeros(cap, dtype=torch.float64, device=dev)
    nnz_off_d = torch.zeros(q + 1, dtype=torch.int64, device=dev)
    dmap.reserve(q, n, m_c, K_NN)

    def step():
        dmap.sample_reject_dev(s_d, g_d, cand_d, q, m_c, clr, n, nodes_d, n_filled_d, n_used_d, stream=stream)
        dmap.build_prm_dev(nodes_d, q, n, K_NN, clr, cdc, row_ptr_d, col_d, w_d, cap, nnz_off_d, stream=stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    if int(n_filled_d.min()) < n:
        raise SystemExit("bench.py: a query ran out of candidates; raise candidates_per_query")

    # ---- timed region: device-resident inputs, CUDA events on the launching stream ----
    dmap.read_lookups(stream)
    dmap.read_filter_counter(stream)
    dmap.read_launches()
    dmap.profile(True)
    dmap.profile_read()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    sync_all()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    stage_ms, stage_calls = dmap.profile_read()
    dmap.profile(False)
    lookups = dmap.read_lookups(stream)
    filt_decided, filt_seen, filt_lookups = dmap.read_filter_counter(stream)
    launches = dmap.read_launches()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    checks_per_step = q * n * K_NN
    value = checks_per_step * world * args.steps / (ms_max * 1e-3)
    nnz_total = int(nnz_off_d[q].item())
    free_frac = None

    # ---- sustained clocks: the same step repeated for >= 2.5 s (the K timed steps above last a fraction of a second
    # at boost clocks); reported next to the headline, not instead of it ----
    sustained = None
    if not args.no_sustained:
        reps = int(min(5000, max(args.steps, 2500.0 / max(ms_max / args.steps, 1e-3))))
        clocks2 = ClockSampler(local)
        if rank == 0:
            clocks2.start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        s0.record(stream)
        for _ in range(reps):
            step()
        s1.record(stream)
        sync_all()
        clk2 = clocks2.stop() if rank == 0 else None
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ms_s = float(ts.item())
        sustained = {"value": checks_per_step * world * reps / (ms_s * 1e-3), "unit": UNIT, "steps": reps,
                     "seconds": ms_s * 1e-3, "ms_per_step": ms_s / reps, "clocks": clk2}
        dmap.read_lookups(stream)
        dmap.read_launches()

    # ---- single-query latency (ms per query, PRM-build part), same device path, q = 1 ----
    def step1():
        dmap.sample_reject_dev(s_d, g_d, cand_d, 1, m_c, clr, n, nodes_d, n_filled_d, n_used_d, stream=stream)
        dmap.build_prm_dev(nodes_d, 1, n, K_NN, clr, cdc, row_ptr_d, col_d, w_d, cap, nnz_off_d, stream=stream)
    single_ms = None
    single_graph_ms = None
    if not args.no_single:
        for _ in range(5):
            step1()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(50):
            step1()
        e1.record(stream)
        torch.cuda.synchronize()
        single_ms = e0.elapsed_time(e1) / 50
        # the same launches replayed from one CUDA graph (no per-launch host cost)
        try:
            side = torch.cuda.Stream()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                cs = torch.cuda.current_stream()
                dmap.sample_reject_dev(s_d, g_d, cand_d, 1, m_c, clr, n, nodes_d, n_filled_d, n_used_d, stream=cs)
                dmap.build_prm_dev(nodes_d, 1, n, K_NN, clr, cdc, row_ptr_d, col_d, w_d, cap, nnz_off_d, stream=cs)
            for _ in range(5):
                graph.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(50):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            single_graph_ms = e0.elapsed_time(e1) / 50
        except Exception as ex:  # capture is an optimisation of the launch path only
            single_graph_ms = None
            print(f"bench.py: CUDA graph capture skipped: {ex}", file=sys.stderr)

    # ---- graph search right after the build: shortest start->goal path of every query, on the device CSR ----
    sssp_leg = None
    if not args.no_sssp:
        src_d = torch.zeros((q, 1), dtype=torch.int32, device=dev)
        dist_d = torch.zeros((q, n), dtype=torch.float64, device=dev)
        pred_d = torch.zeros((q, n), dtype=torch.int32, device=dev)

        def sssp_step():
            dmap.sssp_dev(q, n, row_ptr_d, col_d, w_d, nnz_off_d, src_d, 1, dist_d, pred_d, stream=stream)
        for _ in range(3):
            sssp_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(args.steps, 10))
        e0.record(stream)
        for _ in range(reps):
            sssp_step()
        e1.record(stream)
        torch.cuda.synchronize()
        sssp_ms = e0.elapsed_time(e1) / reps
        reach = float(torch.isfinite(dist_d[:, 1]).float().mean().item())
        # one query alone: the library spreads it over a cluster of 16 CTAs (distributed shared memory, rows staged on chip)
        qi = int(torch.isfinite(dist_d[:, 1]).float().argmax().item())
        off1 = (nnz_off_d[qi:qi + 2] - nnz_off_d[qi]).contiguous()
        lo1, hi1 = int(nnz_off_d[qi].item()), int(nnz_off_d[qi + 1].item())
        col1, w1, rp1 = col_d[lo1:hi1], w_d[lo1:hi1], row_ptr_d[qi:qi + 1]
        d1 = torch.zeros((1, n), dtype=torch.float64, device=dev)
        p1 = torch.zeros((1, n), dtype=torch.int32, device=dev)
        for _ in range(5):
            dmap.sssp_dev(1, n, rp1, col1, w1, off1, src_d[:1], 1, d1, p1, stream=stream)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(50):
            dmap.sssp_dev(1, n, rp1, col1, w1, off1, src_d[:1], 1, d1, p1, stream=stream)
        e1.record(stream)
        torch.cuda.synchronize()
        single_sssp_ms = e0.elapsed_time(e1) / 50
        if not torch.equal(d1[0], dist_d[qi]):
            raise SystemExit("bench.py: single-query shortest paths differ from the batched ones")
        sssp_leg = {"ms_per_step": sssp_ms, "queries": q, "ms_per_query_batched": sssp_ms / q,
                    "single_query_latency_ms": single_sssp_ms, "goal_reachable_fraction": reach,
                    "api": "ctp_sssp_dev, CUDA events (batched: one CTA per query, distances in shared memory; one query "
                           "alone: a cluster of 16 CTAs, distances sliced over distributed shared memory, its rows staged on chip)"}
        if rank == 0 and world == 1 and not args.no_cpu:
            # CPU: the oracle's binary-heap Dijkstra (include/dijkstra.hpp:83-176 restated) on the same CSRs
            from oracle import orc as _orc
            kq = min(q, 16)
            rp_h = row_ptr_d[:kq].cpu().numpy()
            off_h = nnz_off_d[:kq + 1].cpu().numpy()
            col_h = col_d[:int(off_h[kq])].cpu().numpy()
            w_h = w_d[:int(off_h[kq])].cpu().numpy()
            t0 = time.perf_counter()
            mism = 0
            for i in range(kq):
                d0, _ = _orc.sssp(rp_h[i], col_h[off_h[i]:off_h[i + 1]], w_h[off_h[i]:off_h[i + 1]], [0])
                mism += int(not np.array_equal(d0, dist_d[i].cpu().numpy()))
            t_cpu = time.perf_counter() - t0
            sssp_leg["cpu_ms_per_query_single_thread"] = 1e3 * t_cpu / kq
            sssp_leg["distance_mismatch_queries"] = mism

    # ---- clustering-stage batches on the same map: shorten_path and pairwise isDeformablePath ----
    paths_leg = None
    if rank == 0 and not args.no_paths:
        try:
            paths_leg = run_paths_leg(dmap, nodes_d, row_ptr_d, col_d, w_d, nnz_off_d, min(q, 64), n, clr, cdc)
        except Exception as ex:
            print(f"bench.py: paths leg skipped: {ex}", file=sys.stderr)

    # ---- end to end through the host-pointer C ABI (pinned host buffers, copies timed) ----
    # The roadmaps come back in the compact form (CTP_CSR_UPPER | CTP_CSR_NO_WEIGHTS: every undirected edge once, the
    # weights ||to - from|| left to the host-side mirror include/ctopprm/csr_expand.hpp, which reproduces them bit for
    # bit); the same call with the full symmetric weighted CSR is timed next to it, and so is the mirror itself.
    e2e = None
    e2e_full = None
    if not args.no_e2e:
        k2 = args.e2e_steps or max(2, min(args.steps, 5))

        def time_e2e(flags, out, reps):
            for _ in range(2):
                dmap.sample_multiple(s_h, g_h, cand_h, n, clr, cdc, k=K_NN, out=out, csr_flags=flags)
            sync_all()
            dmap.read_io()
            t0 = time.perf_counter()
            for _ in range(reps):
                dmap.sample_multiple(s_h, g_h, cand_h, n, clr, cdc, k=K_NN, out=out, csr_flags=flags)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            io_h2d, io_d2h = dmap.read_io()
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            # bytes actually copied (the library counts them): the candidate streams go up head-first, tails on demand
            return {"value": checks_per_step * world * reps / dt, "unit": UNIT, "h2d_bytes_per_step": int(io_h2d // reps),
                    "d2h_bytes_per_step": int(io_d2h // reps), "ms_per_step": 1e3 * dt / reps, "steps": reps}

        # compact wire form: each undirected edge once, no weights, 16-bit columns where the ids fit, nodes as positions
        # in the candidate streams the host already holds
        c16 = n <= 65536
        compact = capi.CSR_UPPER | capi.CSR_NO_WEIGHTS | capi.NODES_AS_INDEX | (capi.CSR_COL16 if c16 else 0)
        out_c = dict(nodes=capi.pinned_empty((q, n), np.int32), n_filled=capi.pinned_empty((q,), np.int32),
                     row_ptr=capi.pinned_empty((q, n + 1), np.int32),
                     col=capi.pinned_empty((cap // 2,), np.uint16 if c16 else np.int32),
                     nnz_off=capi.pinned_empty((q + 1,), np.int64))
        e2e = time_e2e(compact, out_c, k2)
        e2e["api"] = ("ctp_sample_multiple_ex(CTP_CSR_UPPER | CTP_CSR_NO_WEIGHTS | CTP_NODES_AS_INDEX" + (" | CTP_CSR_COL16" if c16 else "") +
                      "): host candidate streams -> host roadmap in the compact wire form (accepted nodes as positions in their "
                      "candidate streams, each undirected edge once" + (", 16-bit columns" if c16 else "") + "; weights recomputable "
                      "from the nodes), wall clock, max over ranks")
        assert 2 * int(out_c["nnz_off"][q]) == nnz_total, "compact host-API adjacency differs from the device-API CSR"
        if rank == 0 and world == 1:
            # the host-side mirror back to the full symmetric weighted CSR (not part of the timed call: a consumer that
            # walks the upper lists or rebuilds visibility_node_ids does not need it)
            from ctopprm_learn_b200 import hostlib
            thr = cpu_cores()
            full_h = dict(row_ptr=np.empty((q, n + 1), dtype=np.int32), col=np.empty(nnz_total, dtype=np.int32),
                          w=np.empty(nnz_total), nnz_off=np.empty(q + 1, dtype=np.int64))
            t0 = time.perf_counter()
            nodes_h = capi.nodes_from_index(out_c["nodes"], cand_h, s_h, g_h)
            col_h = out_c["col"][:nnz_total // 2].astype(np.int32)
            e2e["host_unpack_ms"] = 1e3 * (time.perf_counter() - t0)  # numpy, one thread (include/ctopprm/csr_expand.hpp has the C++ form)
            hostlib.csr_expand_upper(nodes_h, out_c["row_ptr"], col_h, out_c["nnz_off"], threads=thr, out=full_h)
            t0 = time.perf_counter()
            hostlib.csr_expand_upper(nodes_h, out_c["row_ptr"], col_h, out_c["nnz_off"], threads=thr, out=full_h)
            e2e["host_mirror_ms"] = 1e3 * (time.perf_counter() - t0)
            e2e["host_mirror_threads"] = thr
            same = (np.array_equal(full_h["row_ptr"], row_ptr_d.cpu().numpy())
                    and np.array_equal(full_h["col"], col_d[:nnz_total].cpu().numpy())
                    and np.array_equal(full_h["w"], w_d[:nnz_total].cpu().numpy()))
            e2e["host_mirror_equals_device_csr"] = bool(same)
            del full_h
        if not args.no_e2e_full:
            out_f = dict(nodes=capi.pinned_empty((q, n, 3), np.float64), n_filled=out_c["n_filled"], row_ptr=out_c["row_ptr"],
                         col=capi.pinned_empty((cap,), np.int32), w=capi.pinned_empty((cap,), np.float64),
                         nnz_off=out_c["nnz_off"])
            e2e_full = time_e2e(0, out_f, max(2, k2 // 2))
            e2e_full["api"] = "ctp_sample_multiple: the full symmetric CSR with FP64 weights comes back (round-1 format)"
            assert int(out_f["nnz_off"][q]) == nnz_total, "host-API CSR differs from the device-API CSR"
            del out_f

    # ---- end to end per query: host candidates -> host start->goal paths (PRM build + shortest path) ----
    e2e_query = None
    if not args.no_e2e and not args.no_sssp:
        pcap = 512
        qout = dict(path_ids=capi.pinned_empty((q, pcap), np.int32), path_xyz=capi.pinned_empty((q, pcap, 3), np.float64),
                    path_n=capi.pinned_empty((q,), np.int32), length=capi.pinned_empty((q,), np.float64),
                    n_filled=capi.pinned_empty((q,), np.int32))
        for _ in range(2):
            dmap.query_paths(s_h, g_h, cand_h, n, clr, cdc, k=K_NN, path_cap=pcap, out=qout)
        sync_all()
        k3 = args.e2e_steps or max(2, min(args.steps, 5))
        dmap.read_io()
        t0 = time.perf_counter()
        for _ in range(k3):
            dmap.query_paths(s_h, g_h, cand_h, n, clr, cdc, k=K_NN, path_cap=pcap, out=qout)
        torch.cuda.synchronize()
        dtq = time.perf_counter() - t0
        qio = dmap.read_io()
        ttq = torch.tensor([dtq], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ttq, op=dist.ReduceOp.MAX)
        dtq = float(ttq.item())
        e2e_query = {"ms_per_query": 1e3 * dtq / k3 / q, "ms_per_step": 1e3 * dtq / k3, "queries_per_step_per_gpu": q,
                     "edge_checks_per_s": checks_per_step * world * k3 / dtq,
                     "paths_found_fraction": float((qout["path_n"] > 0).mean()),
                     "h2d_bytes_per_step": int(qio[0] // k3), "d2h_bytes_per_step": int(qio[1] // k3),
                     "api": "ctp_query_paths (host candidate streams -> host start->goal paths; roadmaps stay on the device)"}

    # ---- one query through the reference's class surface (C++ host layer over the C ABI): rank 0, N = 1 ----
    planner_leg = None
    full_query = None
    if rank == 0 and world == 1 and not args.no_planner:
        from ctopprm_learn_b200 import hostlib
        t0 = time.perf_counter()
        hm = hostlib.HostMap.from_memory(vox, c, e, device=local)
        t_load = time.perf_counter() - t0
        qi = 0  # the first query whose goal is reachable, when the query-paths leg has told us
        if e2e_query is not None and (qout["path_n"] > 0).any():
            qi = int(np.argmax(qout["path_n"] > 0))
        ya = hostlib.DEFAULT_YAML.format(map="(memory)", clr=clr, cdc=cdc, planar="false", n=n, s=list(s_h[qi]), g=list(g_h[qi]))
        t_sm, t_sp, found = [], [], 0
        for rep in range(4):
            pl = hostlib.HostPlanner(hm, ya, s_h[qi], g_h[qi])
            pl.set_candidates(cand_h[qi])
            t0 = time.perf_counter()
            pl.sample_multiple(n)      # device PRM build; the CSR comes back, the pointer graph is filled only on demand
            t1 = time.perf_counter()
            ids, length = pl.shortest()  # findShortestPath: ctp_sssp over the same roadmap
            t2 = time.perf_counter()
            pl.close()
            if rep:
                t_sm.append(t1 - t0)
                t_sp.append(t2 - t1)
                found = int(len(ids) > 0)
        planner_leg = {"map_load_ms": 1e3 * t_load, "sample_multiple_ms": 1e3 * float(np.median(t_sm)),
                       "find_shortest_path_ms": 1e3 * float(np.median(t_sp)), "query": qi, "path_found": found,
                       "api": "ESDFMap::loadFromMemory + TopologicalPRMClustering::sampleMultiple / findShortestPath "
                              "(include/ctopprm/*.hpp over the C ABI), wall clock"}
        # the whole CTopPRM query -- PRM build, shortest path, cluster stage (floods, homotopy checks, min tours, search
        # over clusters), shortening, pruning -- through find_geometrical_paths, on the map's face-to-face query
        # (C1-C3 start / goal on opposite faces) and beside the CPU restatement of the same stages
        fs, fg = synth.default_query(c, e, cfg=WORKLOADS[args.workload][0])
        fcand = synth.ellipsoid_candidates(fs, fg, cpn * n, 4242)
        fya = hostlib.DEFAULT_YAML.format(map="(memory)", clr=clr, cdc=cdc, planar="false", n=n, s=list(fs), g=list(fg))
        t_full, lens_g, paths_g = [], None, None
        for rep in range(4):
            t0 = time.perf_counter()
            lens_g, paths_g = hm.find_geometrical_paths(fya, fs, fg, "", cand=fcand, want_paths=True, cap_pts=1 << 20, cap=4096)
            if rep:
                t_full.append(time.perf_counter() - t0)
        full_query = {"ms_per_query": 1e3 * float(np.median(t_full)), "phase_ms_last_call": hostlib.last_phase_times(),
                      "distinct_paths": int(len(lens_g)),
                      "path_lengths": [float(x) for x in lens_g[:8]], "nodes": n,
                      "api": "TopologicalPRMClustering::find_geometrical_paths (sampleMultiple -> findShortestPath -> "
                             "clusterGraph -> shorten -> removeTooLong / removeEquivalent -> shorten -> removeEquivalent), "
                             "host candidates in, host paths out, wall clock, median of 3"}
        if not args.no_cpu:
            from oracle import orc as _orc
            _orc.build()
            om_f = _orc.OracleMap(vox, c, e)
            cores = cpu_cores()
            t0 = time.perf_counter()
            ow = _orc.find_geometrical_paths(om_f, fs, fg, fcand, n, clr, cdc, nthreads=1)
            t_cpu1 = time.perf_counter() - t0
            t0 = time.perf_counter()
            _orc.find_geometrical_paths(om_f, fs, fg, fcand, n, clr, cdc, nthreads=cores)
            t_cpuN = time.perf_counter() - t0
            same = bool(len(ow["lengths"]) == len(lens_g) and np.array_equal(ow["lengths"], lens_g) and
                        all(np.array_equal(a_, b_) for a_, b_ in zip(ow["paths"], paths_g)))
            full_query["cpu_restatement"] = {"ms_per_query_1_thread": 1e3 * t_cpu1, "ms_per_query_all_cores": 1e3 * t_cpuN,
                                             "cores": cores, "kind": "port", "distinct_paths": int(len(ow["lengths"])),
                                             "cluster_seeds": int(len(ow["seeds"])), "capped_refloods": int(ow["capped_refloods"]),
                                             "note": "oracle restatement of the same stages; only the roadmap's directed checks "
                                                     "use more than one thread (the reference is single-threaded throughout)"}
            full_query["paths_identical_to_cpu"] = same
        hm.close()

    # ---- CPU baseline (rank 0, N = 1): bounded sample of the same queries, all host cores ----
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = cpu_cores()
        arm = CpuArm(vox, c, e)
        # single-thread probe on one query, then a sample sized for ~10-20 s of wall time
        t1, o1 = arm.run(s_h[:1], g_h[:1], cand_h[:1], n, clr, cdc, 1)
        # as many of the step's queries as fit in ~20 s of wall time on these cores (all of them if possible)
        q_s = args.cpu_queries or int(min(q, max(cores, cores * 20.0 / max(t1, 1e-3))))
        q_s = max(1, min(q_s, q))
        tc, oc = arm.run(s_h[:q_s], g_h[:q_s], cand_h[:q_s], n, clr, cdc, cores)
        reps = 1
        while tc < 10.0 and reps < 8:  # repeat the sample until it is ~10 s of wall time on all cores
            tc += arm.run(s_h[:q_s], g_h[:q_s], cand_h[:q_s], n, clr, cdc, cores)[0]
            reps += 1
        tc /= reps
        cpu = {"value": q_s * n * K_NN / tc, "unit": UNIT, "cores": cores, "kind": arm.kind,
               "sample": f"first {q_s} of the step's {q} queries x {reps} repeats, whole sampleMultiple on the host "
                         f"({'reference src/esdf_map.cpp + src/base_map.cpp via oracle/_ref' if arm.kind == 'reference' else 'oracle port'}), "
                         f"OpenMP over queries, {tc:.1f} s per repeat",
               "single_thread_value": n * K_NN / t1}
        # the GPU results for the same queries must agree with the CPU arm
        rp = row_ptr_d[:q_s].cpu().numpy()
        gpu_nnz = rp[:, -1].astype(np.int64)
        parity = {"queries_checked": int(q_s), "nnz_mismatch": int((gpu_nnz != oc["nnz"]).sum()),
                  "n_filled_mismatch": int((n_filled_d[:q_s].cpu().numpy() != oc["n_filled"]).sum())}
        free_frac = float(oc["n_free"].sum() / (q_s * n * K_NN))
        if args.parity_queries > 0:
            pq = max(1, min(args.parity_queries, q, 2 if n > 50000 else q))
            parity.update(detailed_parity(dmap, arm.om, nodes_d, n, pq, clr, cdc, dev, torch))
        arm.close()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        edge_ms = stage_ms["edges"]
        launches_edge = max(1, stage_calls["edges"])
        alg_bytes = lookups * 32.0
        achieved = alg_bytes / (edge_ms * 1e-3) / 1e9 if edge_ms > 0 else 0.0
        traffic = lts_bytes = None
        grid_in_hbm = vox.nbytes > (100 << 20)  # beyond the 126 MB L2 the gathers are DRAM traffic
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            ent = tr.get(f"{args.workload}:{dmap.layout_name()}", {})
            traffic = ent.get("dram_bytes_per_launch")
            lts_bytes = ent.get("lts_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, n, q, cpn, world),
            "impl_config": {"layout": dmap.layout_name() + (" (edge march) + cell8 (rejection)" if extra else ""),
                            "lanes_per_edge": dmap.prm_group()},
            "ms_per_query": {"batched": ms_max / args.steps / q, "single_query_latency": single_ms,
                             "single_query_latency_cuda_graph": single_graph_ms, "class_surface": planner_leg,
                             "full_query": full_query},
            "paths": paths_leg, "shortest_path": sssp_leg, "e2e_query": e2e_query,
            "lookups_per_s": lookups / (ms_max * 1e-3),
            "stage_ms_per_step": {k_: v / args.steps for k_, v in stage_ms.items()},
            "roofline": {"bound": "hbm" if grid_in_hbm else "l2", "limiter": "dram gathers" if grid_in_hbm else ("issue (grid L2-resident; float32 whole-edge filter + FP64 march of the rest)" if filt_seen else "issue / fp64 pipe (grid L2-resident)"),
                         "kernel": ("k_edge_mid + k_edge_flat" if filt_seen else "k_edge_flat") if dmap.prm_group() == 1 else "k_edge_march", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "l2_bytes_per_launch": lts_bytes,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes_per_launch": alg_bytes / launches_edge,
                         "lookups_per_launch": lookups / launches_edge, "ms_per_launch": edge_ms / launches_edge,
                         "share_of_step": edge_ms / ms,
                         "whole_edge_filter": {"edges_shown": filt_seen / max(launches_edge, 1), "edges_decided_fraction": filt_decided / max(filt_seen, 1),
                                               "lookups_performed_per_launch": (filt_seen + lookups - filt_lookups) / max(launches_edge, 1),
                                               "note": "achieved / algorithmic bytes count the reference-order lookups (what the CPU performs); the "
                                                       "float32 whole-edge filter decides most free roadmap edges from one lookup, with identical verdicts"}},
            "e2e": e2e, "e2e_full_csr": e2e_full, "numa_binding_rank0": numa, "sustained": sustained, "gpu_launches": int(launches), "clocks": clk,
            "cpu_baseline": cpu, "csr_nnz_per_step": nnz_total, "free_edge_fraction": free_frac, "parity": parity,
            "reference_note": "the CPU arm's neighbour stage is the oracle's exact float32 k-NN (FLANN is absent from the "
                              "reference tree); its collision code is the reference's own src/esdf_map.cpp + src/base_map.cpp",
        }
        if world == 1 and args.workload == "C2" and not args.no_legs:
            line["legs"] = run_legs(args)
        emit_line(line)
    dmap.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
