#!/usr/bin/env python
"""Aggregate an ncu --csv launch list (gpu__time_duration.sum [+ other metrics]) per kernel name.
usage: tools/launch_summary.py launches.csv [skip_first_n_launches_per_kernel]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = None
per = collections.OrderedDict()
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0]
        per.setdefault(name, collections.OrderedDict()).setdefault(d["ID"], {})[d["Metric Name"]] = (
            float(d["Metric Value"].replace(",", "")), d["Metric Unit"])
tot = 0.0
out = []
for name, launches in per.items():
    ls = list(launches.values())[skip:]
    if not ls:
        continue
    t = 0.0
    for m in ls:
        v, u = m["gpu__time_duration.sum"]
        t += v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    extra = ""
    if "smsp__inst_executed.sum" in ls[0]:
        inst = sum(m["smsp__inst_executed.sum"][0] for m in ls) / len(ls)
        act = sum(m["smsp__thread_inst_executed_per_inst_executed.ratio"][0] for m in ls) / len(ls)
        extra = f" warp-inst/launch={inst:.3g} lanes/inst={act:.1f}"
    out.append((t, name, len(ls), extra))
    tot += t
for t, name, cnt, extra in sorted(out, reverse=True):
    print(f"{name[:64]:64s} n={cnt:3d} total={t:9.3f} ms avg={t / cnt:8.4f} ms share={t / tot:.3f}{extra}")
print(f"{'TOTAL':64s} {tot:.3f} ms")
