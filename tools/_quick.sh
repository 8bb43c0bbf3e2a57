python -m pytest tests/test_gpu_parity.py -x -q -k "csr or build_prm or adjacency or compact or c2_full or randomized or small_k or missing or sample_multiple" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-single --no-sssp --no-planner --no-paths --no-sustained --no-legs 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2', round(d['value']/1e9,3), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, round(d['roofline']['frac'],3))"
