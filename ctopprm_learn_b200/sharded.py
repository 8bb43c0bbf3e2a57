"""Multi-GPU sharding of the PRM-build path over torch.distributed (one process per GPU).

Two regimes:

* independent queries / maps (BASELINE configs[3], [4]): ``shard_range`` gives each rank a contiguous
  block of queries; the map is replicated; no collective on the data path.
* one large query (configs[2]): ``sample_multiple_sharded`` keeps everything in device memory:

    1. rejection of this rank's contiguous block of the candidate stream (``ctp_sample_reject_dev``);
    2. all-gather of the per-rank accept counts (order-preserving acceptance only needs their prefix);
    3. **all-gather of the accepted node sets** (padded blocks, rank order = stream order) -- every rank now holds the
       same n nodes;
    4. neighbour search + directed edge checks for this rank's slice of the nodes (``ctp_prm_rows_dev``: a range of
       the k-NN grid's cell order, i.e. a compact region of space, aligned to cell boundaries so that all ranks agree);
    5. element-wise MAX all-reduce of the neighbour table and of the verdict bytes (rows a rank does not own hold a
       value below every valid entry);
    6. symmetric union -> CSR on every rank (``ctp_csr_from_directed_dev``; a few percent of the work).

  The result is identical to the single-GPU build.  Collectives are NCCL on GPUs; the CPU tests run the same code
  over gloo with the test oracle injected as the backend (no device exists there).
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist

NBR_UNOWNED = np.int32(-2139062144)  # 0x80808080: what ctp_prm_rows_dev leaves in rows outside the slice


def shard_range(total: int, rank: int, world: int):
    """contiguous block [lo, hi) of `total` items for `rank`; sizes differ by at most one"""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _device_for(backend_name: str, local_device):
    return torch.device("cuda", local_device) if backend_name == "nccl" else torch.device("cpu")


class DeviceBackend:
    """the product backend: every operation is a C-ABI call on this rank's GPU, on torch CUDA tensors"""

    def __init__(self, dmap):
        self.m = dmap
        self.dev = torch.device("cuda", dmap.device)

    def to_device(self, a, dtype):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.dev, non_blocking=True)

    def accept_block(self, start_t, goal_t, cand_t, clr, cap):
        """first `cap` accepted candidates of a block, in stream order -> (cap x 3 tensor, zero padded; count tensor)"""
        st = torch.cuda.current_stream(self.dev)
        nodes = torch.zeros((cap + 2, 3), dtype=torch.float64, device=self.dev)
        nf = torch.zeros(1, dtype=torch.int32, device=self.dev)
        nu = torch.zeros(1, dtype=torch.int32, device=self.dev)
        if cand_t.shape[0] > 0 and cap > 0:
            self.m.sample_reject_dev(start_t, goal_t, cand_t, 1, cand_t.shape[0], clr, cap + 2, nodes, nf, nu, stream=st)
        else:
            nf.fill_(2)
        return nodes[2:], (nf - 2).to(torch.int64)

    def prm_rows(self, nodes_t, k, clr, cdc, first, last):
        n = nodes_t.shape[0]
        st = torch.cuda.current_stream(self.dev)
        nbr = torch.empty((n, k), dtype=torch.int32, device=self.dev)
        free = torch.empty((n, k), dtype=torch.uint8, device=self.dev)
        self.m.prm_rows_dev(nodes_t, n, k, clr, cdc, first, last, nbr, free, stream=st)
        return nbr, free

    def csr(self, nodes_t, nbr_t, free_t):
        n, k = nbr_t.shape
        st = torch.cuda.current_stream(self.dev)
        cap = 2 * n * k
        rp = torch.empty(n + 1, dtype=torch.int32, device=self.dev)
        col = torch.empty(cap, dtype=torch.int32, device=self.dev)
        w = torch.empty(cap, dtype=torch.float64, device=self.dev)
        off = torch.zeros(2, dtype=torch.int64, device=self.dev)
        self.m.csr_from_directed_dev(nodes_t, 1, n, k, nbr_t, free_t, rp, col, w, cap, off, stream=st)
        return rp, col, w, off

    def sync(self):
        torch.cuda.synchronize(self.dev)


def sample_multiple_sharded(backend, start, goal, cand, n, clr, cdc, k=13, local_device=0, timings=None, to_host=True):
    """TopologicalPRMClustering::sampleMultiple (include/topological_prm_clustering.hpp:288-401) for ONE query with the
    work split over all ranks of the default process group.  `cand`: the full candidate stream (identical on every
    rank, e.g. read from the shared sample file; each rank uploads only its own block) or, as a torch tensor already
    on the device, this rank's block.  Returns, on every rank, dict(nodes, nbr, directed_free, row_ptr, col, w)
    identical to the single-GPU result (numpy when to_host, else device tensors with nnz).  timings (optional dict):
    seconds per phase on this rank, measured after a synchronisation of the device."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _device_for(dist.get_backend(), local_device)

    def tick(name, t0):
        if timings is not None:
            backend.sync()
            t1 = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + (t1 - t0)
            return t1
        return t0

    start_t = backend.to_device(np.asarray(start, dtype=np.float64).reshape(1, 3), torch.float64)
    goal_t = backend.to_device(np.asarray(goal, dtype=np.float64).reshape(1, 3), torch.float64)
    if isinstance(cand, torch.Tensor):
        block = cand
    else:
        cand = np.ascontiguousarray(cand, dtype=np.float64).reshape(-1, 3)
        lo, hi = shard_range(len(cand), rank, world)
        block = backend.to_device(cand[lo:hi], torch.float64)
    backend.sync()
    t0 = time.perf_counter()
    # 1. rejection on this rank's contiguous block of the stream; the first n-2 accepted overall win
    cap = n - 2
    mine, count = backend.accept_block(start_t, goal_t, block, clr, min(cap, max(int(block.shape[0]), 0)))
    if mine.shape[0] < cap:  # equal block shapes for the gather
        mine = torch.cat([mine, torch.zeros((cap - mine.shape[0], 3), dtype=torch.float64, device=dev)])
    t0 = tick("reject", t0)
    # 2 + 3. accept counts, then the node sets (rank order == stream order)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, count.reshape(1))
    gathered = torch.empty((world * cap, 3), dtype=torch.float64, device=dev)  # concatenated form: gloo accepts only this
    dist.all_gather_into_tensor(gathered, mine.contiguous())
    gathered = gathered.view(world, cap, 3)
    cnt = [int(c) for c in counts.cpu()]  # the one host read of the exchange: how many nodes each block contributed
    parts = [start_t, goal_t] + [gathered[r, :cnt[r]] for r in range(world)]
    nodes_t = torch.cat(parts)[:n].contiguous()
    if nodes_t.shape[0] < n:
        raise RuntimeError(f"candidate stream exhausted: {nodes_t.shape[0]} of {n} nodes accepted")
    t0 = tick("allgather_nodes", t0)
    # 4. this rank's slice of the neighbour search and of the directed checks
    rlo, rhi = shard_range(n, rank, world)
    nbr_t, free_t = backend.prm_rows(nodes_t, k, clr, cdc, rlo, rhi)
    t0 = tick("knn_and_edges", t0)
    # 5. assemble the tables: rows a rank does not own are below every valid value
    if world > 1:
        dist.all_reduce(nbr_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(free_t, op=dist.ReduceOp.MAX)
    t0 = tick("allreduce_tables", t0)
    # 6. symmetric union -> CSR (replicated)
    rp, col, w, off = backend.csr(nodes_t, nbr_t, free_t)
    t0 = tick("csr", t0)
    if not to_host:
        return dict(nodes=nodes_t, nbr=nbr_t, directed_free=free_t, row_ptr=rp, col=col, w=w, nnz_off=off)
    nnz = int(off[1])
    return dict(nodes=nodes_t.cpu().numpy(), nbr=nbr_t.cpu().numpy(), directed_free=free_t.cpu().numpy().astype(bool),
                row_ptr=rp.cpu().numpy(), col=col[:nnz].cpu().numpy(), w=w[:nnz].cpu().numpy())


def gather_query_results(local: dict, local_device=0):
    """batched-queries regime: every rank built CSRs for its own block of queries; collect the per-query
    edge counts on all ranks (the only cross-rank traffic of that regime besides timing)"""
    dev = _device_for(dist.get_backend(), local_device)
    world = dist.get_world_size()
    nnz = torch.as_tensor(np.ascontiguousarray(np.diff(local["nnz_off"]).astype(np.int64))).to(dev)
    cnt = torch.tensor([nnz.shape[0]], dtype=torch.int64, device=dev)
    cnts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cnts, cnt)
    cnts = [int(c) for c in cnts.cpu()]
    mx = max(max(cnts), 1)
    pad = torch.zeros(mx, dtype=torch.int64, device=dev)
    pad[:nnz.shape[0]] = nnz
    out = torch.empty(world * mx, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, pad)
    out = out.view(world, mx)
    return np.concatenate([out[r, :cnts[r]].cpu().numpy() for r in range(world)])
