"""Multi-GPU sharding of the PRM-build path over torch.distributed (one process per GPU).

Two regimes:

* independent queries / maps (BASELINE configs[3], [4]): ``shard_range`` gives each rank a contiguous
  block of queries; the map is replicated; no collective on the data path.
* one large query (configs[2]): ``sample_multiple_sharded`` splits the candidate stream into contiguous
  blocks (order-preserving acceptance needs only the per-rank accept counts), all-gathers the accepted
  nodes, lets every rank run the directed edge checks for its own range of source rows, all-gathers the
  verdict bytes and builds the symmetric CSR -- the result is identical to the single-GPU build.

The compute itself is the CUDA path behind the C ABI (``capi.DeviceMap``); this module only decides who
does what and moves the small pieces between ranks (NCCL on GPUs; gloo in the CPU tests, which inject
the test oracle as the backend because no device exists there).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """contiguous block [lo, hi) of `total` items for `rank`; sizes differ by at most one"""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _device_for(backend_name: str, local_device):
    return torch.device("cuda", local_device) if backend_name == "nccl" else torch.device("cpu")


def _all_gather_var(arr: np.ndarray, dev) -> list:
    """all_gather of per-rank arrays whose first dimension differs (padded to the maximum)"""
    world = dist.get_world_size()
    cnt = torch.tensor([arr.shape[0]], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = [int(c.item()) for c in cnts]
    mx = max(max(cnts), 1)
    pad = np.zeros((mx,) + arr.shape[1:], dtype=arr.dtype)
    pad[:arr.shape[0]] = arr
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return [o.cpu().numpy()[:c] for o, c in zip(outs, cnts)]


class DeviceBackend:
    """the product backend: every operation is a C-ABI call on this rank's GPU"""

    def __init__(self, dmap):
        self.m = dmap

    def accepted(self, start, goal, cand, clr, cap):
        nodes, n_filled, _, _ = self.m.sample_reject(start, goal, cand[None], clr, cap + 2, allow_short=True)
        return nodes[0, 2:int(n_filled[0])]

    def knn(self, nodes, k):
        return self.m.knn(nodes, k)[0][0]

    def edge_rows(self, nodes, nbr, lo, hi, clr, cdc):
        k = nbr.shape[1]
        frm = np.repeat(np.arange(lo, hi, dtype=np.int32), k)
        to = np.ascontiguousarray(nbr[lo:hi].reshape(-1))
        free, _, _ = self.m.edge_check_indexed(nodes, frm, to, clr, cdc)
        return free.reshape(hi - lo, k).astype(np.uint8)

    def csr(self, nodes, nbr, directed_free):
        out = self.m.csr_from_directed(nodes, nbr, directed_free)
        return out["row_ptr"][0], out["col"], out["w"]


def sample_multiple_sharded(backend, start, goal, cand, n, clr, cdc, k=13, local_device=0):
    """TopologicalPRMClustering::sampleMultiple (include/topological_prm_clustering.hpp:288-401) for ONE
    query with the work split over all ranks of the default process group.  `cand` is the full candidate
    stream (identical on every rank, e.g. read from the shared sample file).  Returns, on every rank,
    dict(nodes, nbr, directed_free, row_ptr, col, w) identical to the single-GPU result."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _device_for(dist.get_backend(), local_device)
    cand = np.ascontiguousarray(cand, dtype=np.float64).reshape(-1, 3)
    start = np.asarray(start, dtype=np.float64)
    goal = np.asarray(goal, dtype=np.float64)
    # 1. rejection on this rank's contiguous block of the stream; the first n-2 accepted overall win
    lo, hi = shard_range(len(cand), rank, world)
    mine = backend.accepted(start, goal, cand[lo:hi], clr, min(n - 2, hi - lo)) if hi > lo else np.zeros((0, 3))
    parts = _all_gather_var(np.ascontiguousarray(mine), dev)  # rank order == stream order
    acc = np.concatenate(parts) if parts else np.zeros((0, 3))
    if len(acc) < n - 2:
        raise RuntimeError(f"candidate stream exhausted: {len(acc) + 2} of {n} nodes accepted")
    nodes = np.concatenate([start[None], goal[None], acc[:n - 2]])
    # 2. neighbour lists (cheap, replicated), then the directed checks of this rank's source rows
    nbr = backend.knn(nodes, k)
    rlo, rhi = shard_range(n, rank, world)
    mine_free = backend.edge_rows(nodes, nbr, rlo, rhi, clr, cdc) if rhi > rlo else np.zeros((0, k), dtype=np.uint8)
    directed_free = np.concatenate(_all_gather_var(np.ascontiguousarray(mine_free), dev))
    # 3. symmetric union -> CSR (replicated; it is a few percent of the edge checks)
    row_ptr, col, w = backend.csr(nodes, nbr, directed_free)
    return dict(nodes=nodes, nbr=nbr, directed_free=directed_free.astype(bool), row_ptr=row_ptr, col=col, w=w)


def gather_query_results(local: dict, local_device=0):
    """batched-queries regime: every rank built CSRs for its own block of queries; collect the per-query
    edge counts on all ranks (the only cross-rank traffic of that regime besides timing)"""
    dev = _device_for(dist.get_backend(), local_device)
    nnz = np.ascontiguousarray(np.diff(local["nnz_off"]).astype(np.int64))
    return np.concatenate(_all_gather_var(nnz, dev))
