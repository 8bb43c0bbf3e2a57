for A in "--seg-bin 0 --layout tex" "--seg-bin 1 --layout tex" "--seg-bin 0 --layout cell8" "--seg-bin 1 --layout cell8" "--seg-bin 1 --layout brick" "" ; do python bench.py --workload MU --mu-map C3 --steps 10 --no-cpu --no-e2e $A 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MU C3', '$A', round(d['value']/1e9,3), 'G seg/s', round(d['ms_per_step'],3), 'ms', round(d['roofline']['frac'],3), d['lookups_per_segment'])"; done
python bench.py --workload MU --mu-map C2 --steps 10 --no-cpu --no-e2e --seg-bin 1 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MU C2 bin', round(d['value']/1e9,3), 'G seg/s', round(d['ms_per_step'],3), 'ms', round(d['roofline']['frac'],3))"
