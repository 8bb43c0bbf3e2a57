python -m pytest tests/test_gpu_cluster.py tests/test_host_layer.py -x -q -m gpu 2>&1 | tail -3
for W in C2 C3; do python bench.py --workload $W --steps 5 --no-cpu --no-e2e --no-single --no-sssp --no-paths --no-sustained --no-legs 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); fq=d['ms_per_query']['full_query']; print('$W', fq['ms_per_query'], fq['distinct_paths'], fq['phase_ms_last_call'])"; done
