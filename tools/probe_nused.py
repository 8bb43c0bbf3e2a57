"""probe: how many candidates does the rejection loop consume per query on the bench workload?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ctopprm_learn_b200 import capi, synth
import bench
vox, c, e = synth.make_config_map("C2")
m = capi.DeviceMap(vox, c, e)
n = 10000
s, g, cand = bench.make_queries(c, e, 512, n, 4, 0, 128)
nodes, nf, nu, acc = m.sample_reject(s, g, cand, 0.2, n)
r = nu / n
print("n_used / n: min %.2f  p10 %.2f  median %.2f  p90 %.2f  p99 %.2f  max %.2f" % (r.min(), *np.percentile(r, [10, 50, 90, 99]), r.max()))
for f in (1.5, 2.0, 2.5, 3.0):
    print(f"queries needing more than {f} n candidates: {(r > f).mean() * 100:.1f} %")
