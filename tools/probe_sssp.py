"""probe: latency of one shortest-path search (q = 1) and of the PRM build before it, C2 sizes"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ctopprm_learn_b200 import capi, synth

vox, c, e = synth.make_config_map("C2")
m = capi.DeviceMap(vox, c, e, layouts=capi.LAYOUT_PLAIN | capi.LAYOUT_TEX)
m.set_layout(capi.LAYOUT_TEX)
n, k = 10000, 13
S, G = synth.random_queries(c, e, 200, 4)
ok = (m.clearance(S) > 0.3) & (m.clearance(G) > 0.3)
s, g = S[ok][0], G[ok][0]
cand = synth.ellipsoid_candidates(s, g, 4 * n, 1)
dev = torch.device("cuda")
st = torch.cuda.current_stream()
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
s_d, g_d, c_d = T(s[None]), T(g[None]), T(cand[None])
nodes = torch.zeros((1, n, 3), dtype=torch.float64, device=dev)
nf = torch.zeros(1, dtype=torch.int32, device=dev); nu = torch.zeros(1, dtype=torch.int32, device=dev)
rp = torch.zeros((1, n + 1), dtype=torch.int32, device=dev)
cap = 2 * n * k
col = torch.zeros(cap, dtype=torch.int32, device=dev); w = torch.zeros(cap, dtype=torch.float64, device=dev)
off = torch.zeros(2, dtype=torch.int64, device=dev)
src = torch.zeros((1, 1), dtype=torch.int32, device=dev)
dist = torch.zeros((1, n), dtype=torch.float64, device=dev); pred = torch.zeros((1, n), dtype=torch.int32, device=dev)
m.reserve(1, n, 4 * n, k)
def build():
    m.sample_reject_dev(s_d, g_d, c_d, 1, 4 * n, 0.2, n, nodes, nf, nu, stream=st)
    m.build_prm_dev(nodes, 1, n, k, 0.2, 0.05, rp, col, w, cap, off, stream=st)
def search():
    m.sssp_dev(1, n, rp, col, w, off, src, 1, dist, pred, stream=st)
def search_one_cta():
    m.set_tunable(capi.TUNE_SSSP_CLUSTER, 0)
    search()
    m.set_tunable(capi.TUNE_SSSP_CLUSTER, -1)
build(); torch.cuda.synchronize()
search(); d_cluster = dist.clone(); search_one_cta(); torch.cuda.synchronize()
print("cluster == one CTA:", bool(torch.equal(d_cluster, dist)))
for f, name in ((build, "PRM build"), (search, "shortest path (cluster of 16 CTAs, rows staged on chip)"), (search_one_cta, "shortest path (one CTA)")):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): f()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 50:.3f} ms per query (q = 1)")
print("goal distance", float(dist[0, 1]), "straight line", float(np.linalg.norm(g - s)))
