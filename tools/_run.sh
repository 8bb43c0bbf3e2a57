python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -3
python bench.py --steps 10 --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err; tail -3 gpurun_out/b.err; python tools/bsum.py sssp < gpurun_out/b.json; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print(json.dumps({k:d[k] for k in ('ms_per_query','shortest_path','cpu_baseline')}, indent=1))"
