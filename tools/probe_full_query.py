"""probe: one CTopPRM query (find_geometrical_paths) on the C2 map, a few repetitions -- for an ncu launch list"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ctopprm_learn_b200 import hostlib, synth

n, cpn = 10000, 4
vox, c, e = synth.make_config_map("C2")
hm = hostlib.HostMap.from_memory(vox, c, e, device=0)
fs, fg = synth.default_query(c, e, cfg="C2")
cand = synth.ellipsoid_candidates(fs, fg, cpn * n, 4242)
ya = hostlib.DEFAULT_YAML.format(map="(memory)", clr=synth.MIN_CLEARANCE, cdc=synth.COLLISION_DISTANCE_CHECK, planar="false", n=n,
                                 s=list(fs), g=list(fg))
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    t0 = time.perf_counter()
    lens, paths = hm.find_geometrical_paths(ya, fs, fg, "", cand=cand, want_paths=True, cap_pts=1 << 20, cap=4096)
    print("query ms", 1e3 * (time.perf_counter() - t0), "paths", len(lens), hostlib.last_phase_times())
