python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -3
for O in 0 1; do for L in tex plain; do python bench.py --steps 10 --layout $L --edge-order $O --no-cpu --no-e2e --no-single --no-paths --no-sssp | python tools/bsum.py order${O}_$L; done; done
for O in 0 1; do python bench.py --workload C3 --steps 5 --edge-order $O --no-cpu --no-e2e --no-single --no-paths --no-sssp | python tools/bsum.py C3_order$O; done
