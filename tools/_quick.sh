cp ctopprm_learn_b200/lib/libctopprm_b200.so /tmp/keep.so
for R in 2 8; do cp tools/_variants/lib_r$R.so ctopprm_learn_b200/lib/libctopprm_b200.so
for A in "--mu-map C3" "--mu-map C2"; do python bench.py --workload MU --steps 10 --no-cpu --no-e2e $A 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MU R0=$R', '$A', round(d['value']/1e9,3), 'G seg/s', round(d['ms_per_step'],3), 'ms', round(d['roofline']['frac'],3))"; done; done
cp /tmp/keep.so ctopprm_learn_b200/lib/libctopprm_b200.so
