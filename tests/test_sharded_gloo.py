"""world_size-2 (and 3) gloo runs of the sharding logic on CPU.  There is no device here, so the test
injects the ORACLE as the compute backend; what is under test is the host logic of
ctopprm_learn_b200/sharded.py: block assignment, order-preserving merge of accepted nodes, row-range
slices of the neighbour search and of the edge checks, the element-wise MAX merge of the tables -- over the same
collectives the GPUs use (all_gather_into_tensor, all_reduce MAX); the result must equal the unsharded build."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ctopprm_learn_b200 import sharded, synth

CLR, CDC = synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 4096, 100003):
        for world in (1, 2, 3, 8):
            edges = [sharded.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


class OracleBackend:
    """test-only stand-in for sharded.DeviceBackend: the same tensor interface, computed by the oracle on the CPU"""

    def __init__(self, vox, c, e):
        from oracle import orc
        self.orc = orc
        self.om = orc.OracleMap(vox, c, e)

    def to_device(self, a, dtype):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype)

    def sync(self):
        pass

    def accept_block(self, start_t, goal_t, cand_t, clr, cap):
        out = torch.zeros((cap, 3), dtype=torch.float64)
        if cand_t.shape[0] == 0 or cap == 0:
            return out, torch.zeros(1, dtype=torch.int64)
        nodes, filled, _, _ = self.om.sample_reject(start_t.numpy()[0], goal_t.numpy()[0], cand_t.numpy(), clr, cap + 2)
        out[:filled - 2] = torch.from_numpy(nodes[2:filled])
        return out, torch.tensor([filled - 2], dtype=torch.int64)

    def prm_rows(self, nodes_t, k, clr, cdc, first, last):
        nodes = nodes_t.numpy()
        n = len(nodes)
        nbr = np.full((n, k), sharded.NBR_UNOWNED, dtype=np.int32)
        free = np.zeros((n, k), dtype=np.uint8)
        if last > first:
            nbr[first:last] = self.orc.knn_grid(nodes, k)[0][first:last]
            frm = np.repeat(np.arange(first, last, dtype=np.int32), k)
            to = np.ascontiguousarray(nbr[first:last].reshape(-1))
            f, _, _ = self.om.edge_index(nodes, frm, np.maximum(to, 0), clr, cdc)
            free[first:last] = (f & (to >= 0)).reshape(last - first, k)
        return torch.from_numpy(nbr), torch.from_numpy(free)

    def csr(self, nodes_t, nbr_t, free_t):
        import ctypes as C
        n, k = nbr_t.shape
        row_ptr = np.empty(n + 1, dtype=np.int32)
        col = np.empty(2 * n * k, dtype=np.int32)
        w = np.empty(2 * n * k)
        L = self.orc._L()
        L.orc_csr_from_directed.restype = C.c_int64
        nn = np.ascontiguousarray(nodes_t.numpy())
        nb = np.ascontiguousarray(nbr_t.numpy(), dtype=np.int32)
        fr = np.ascontiguousarray(free_t.numpy(), dtype=np.uint8)
        nnz = L.orc_csr_from_directed(nn.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int(k), nb.ctypes.data_as(C.c_void_p),
                                      fr.ctypes.data_as(C.c_void_p), row_ptr.ctypes.data_as(C.c_void_p),
                                      col.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p))
        return (torch.from_numpy(row_ptr), torch.from_numpy(col), torch.from_numpy(w),
                torch.tensor([0, nnz], dtype=torch.int64))


def _worker(rank, world, port, n, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vox, c, e = synth.make_config_map("C1")
        be = OracleBackend(vox, c, e)
        s, g = synth.default_query(c, e)
        cand = synth.ellipsoid_candidates(s, g, 5 * n, 3)
        out = sharded.sample_multiple_sharded(be, s, g, cand, n, CLR, CDC)
        want_nodes = be.om.sample_reject(s, g, cand, CLR, n)[0]
        want = be.om.build_prm(want_nodes, CLR, CDC)
        ok = (np.array_equal(out["nodes"], want_nodes) and np.array_equal(out["nbr"], want["nbr"])
              and np.array_equal(out["directed_free"], want["directed_free"]) and np.array_equal(out["row_ptr"], want["row_ptr"])
              and np.array_equal(out["col"], want["col"]) and np.array_equal(out["w"], want["w"]))
        # batched-queries regime: each rank owns a block of queries, only the edge counts travel
        lo, hi = sharded.shard_range(5, rank, world)
        local = dict(nnz_off=np.concatenate([[0], np.cumsum(np.arange(lo, hi) + 10)]))
        allnnz = sharded.gather_query_results(local)
        ok = ok and np.array_equal(allnnz, np.arange(5) + 10)
        # a stream that is too short is an error on every rank
        try:
            sharded.sample_multiple_sharded(be, s, g, cand[:n // 2], n, CLR, CDC)
            ok = False
        except RuntimeError:
            pass
        open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_single_query_equals_unsharded(world, tmp_path, orc):
    mp.spawn(_worker, args=(world, _free_port(), 700, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"ok{r}").read() == "1"
