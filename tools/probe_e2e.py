"""probe: where the end-to-end call spends its time -- strided (2-D) vs contiguous host->device copies of the candidate
heads, and ctp_sample_multiple_ex at several chunk sizes / with the compute alone"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cuda import cudart
from ctopprm_learn_b200 import capi, synth

q, cnt, n = 512, 40000, 10000
head = int(2.3 * n) + 256
host = capi.pinned_empty((q, cnt, 3), np.float64)
host[:] = 1.0
dev = torch.empty((q, cnt, 3), dtype=torch.float64, device="cuda")
st = torch.cuda.Stream()
sp = st.cuda_stream
def t_copy(label, fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{label:46s} {nbytes / dt / 1e9:6.1f} GB/s  {dt * 1e3:6.2f} ms")
hp, dp = host.ctypes.data, dev.data_ptr()
K = cudart.cudaMemcpyKind.cudaMemcpyHostToDevice
t_copy("1-D copy of the whole stream block", lambda: cudart.cudaMemcpyAsync(dp, hp, q * cnt * 24, K, sp), q * cnt * 24)
t_copy("2-D copy, head of every stream (one call)", lambda: cudart.cudaMemcpy2DAsync(dp, cnt * 24, hp, cnt * 24, head * 24, q, K, sp), q * head * 24)
def chunks(qc):
    for q0 in range(0, q, qc):
        cudart.cudaMemcpy2DAsync(dp + q0 * cnt * 24, cnt * 24, hp + q0 * cnt * 24, cnt * 24, head * 24, min(qc, q - q0), K, sp)
t_copy("2-D copy, 64 streams per call", lambda: chunks(64), q * head * 24)
def rows():
    for i in range(q):
        cudart.cudaMemcpyAsync(dp + i * cnt * 24, hp + i * cnt * 24, head * 24, K, sp)
t_copy("1-D copy per stream head (512 calls)", rows, q * head * 24)
