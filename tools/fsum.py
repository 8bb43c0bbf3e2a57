#!/usr/bin/env python
"""one-line summary of the whole-edge filter figures of a bench.py JSON line read from stdin"""
import json, sys
for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line); r = d["roofline"]; w = r.get("whole_edge_filter", {})
        print(sys.argv[1] if len(sys.argv) > 1 else "", "edges_ms", round(d["stage_ms_per_step"]["edges"], 3), "decided", round(w.get("edges_decided_fraction", 0), 3),
              "lookups alg/performed", int(r["lookups_per_launch"]), int(w.get("lookups_performed_per_launch", 0)), "step", round(d["ms_per_step"], 3), "frac", round(r["frac"], 3))
