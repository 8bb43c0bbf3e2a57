#!/usr/bin/env python
"""print a one-line summary of a bench.py JSON line read from stdin: tools/bsum.py LABEL"""
import json
import sys
label = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    st = {k: round(v, 3) for k, v in d.get("stage_ms_per_step", {}).items()}
    e2e = d.get("e2e") or {}
    print(label, f"value={d['value']:.4g}", f"ms/step={d['ms_per_step']:.3f}", st,
          f"roof={d.get('roofline', {}).get('frac', 0):.3f}", f"e2e={e2e.get('value', 0):.4g}",
          f"single_ms={(d.get('ms_per_query') or {}).get('single_query_latency') if isinstance(d.get('ms_per_query'), dict) else None}")
