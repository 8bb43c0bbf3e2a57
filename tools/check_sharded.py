"""torchrun script: single-query PRM build sharded over the ranks (NCCL) must equal the one-GPU build.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctopprm_learn_b200 import capi, sharded, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
vox, c, e = synth.make_config_map("C2")
m = capi.DeviceMap(vox, c, e, device=local, layouts=capi.LAYOUT_PLAIN | capi.LAYOUT_TEX)
m.set_layout(capi.LAYOUT_TEX)
clr, cdc, n = synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK, 100000
s, g = synth.default_query(c, e)
cand = synth.ellipsoid_candidates(s, g, 4 * n, 77)
be = sharded.DeviceBackend(m)
sharded.sample_multiple_sharded(be, s, g, cand, 2000, clr, cdc, local_device=local)  # warm-up
dist.barrier()
t0 = time.perf_counter()
out = sharded.sample_multiple_sharded(be, s, g, cand, n, clr, cdc, local_device=local)
dist.barrier()
t1 = time.perf_counter()
one = m.sample_multiple(s[None], g[None], cand[None], n, clr, cdc)
nnz = int(one["nnz_off"][1])
ok = (np.array_equal(out["nodes"], one["nodes"][0]) and np.array_equal(out["row_ptr"], one["row_ptr"][0])
      and np.array_equal(out["col"], one["col"][:nnz]) and np.array_equal(out["w"], one["w"][:nnz]))
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded single query over {world} ranks: n={n} nnz={nnz} identical_to_one_gpu={bool(flag.item())} "
          f"wall={1e3 * (t1 - t0):.1f} ms (host-pointer API, numpy staging)")
m.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
