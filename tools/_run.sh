python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -3
python bench.py --steps 10 --no-cpu --no-e2e --no-single --no-paths | python tools/bsum.py cleaned
