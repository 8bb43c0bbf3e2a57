"""probe: timeline of one ctp_sample_multiple_ex call (CTP_PIPE_TRACE=1) on the C2 batch"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ctopprm_learn_b200 import capi, synth
import bench
q, n, cpn = 512, 10000, 4  # candidates per node
vox, c, e = synth.make_config_map("C2")
m = capi.DeviceMap(vox, c, e, layouts=capi.LAYOUT_TEX | capi.LAYOUT_PLAIN)
m.set_layout(capi.LAYOUT_TEX)
cand = capi.pinned_empty((q, cpn * n, 3), np.float64)
s, g, _ = bench.make_queries(c, e, q, n, cpn, 0, q, out=cand, cfg="C2")
cap = q * n * 13
out = dict(nodes=capi.pinned_empty((q, n, 3), np.float64), n_filled=capi.pinned_empty((q,), np.int32),
           row_ptr=capi.pinned_empty((q, n + 1), np.int32), col=capi.pinned_empty((cap,), np.int32),
           nnz_off=capi.pinned_empty((q + 1,), np.int64))
flags = capi.CSR_UPPER | capi.CSR_NO_WEIGHTS
for _ in range(3):
    m.sample_multiple(s, g, cand, n, synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK, k=13, out=out, csr_flags=flags)
os.environ["CTP_PIPE_TRACE"] = "1"
t0 = time.perf_counter()
m.sample_multiple(s, g, cand, n, synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK, k=13, out=out, csr_flags=flags)
print("wall ms", 1e3 * (time.perf_counter() - t0))
