python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -3
for PC in 16 20 24 28; do CTP_DEBUG_KNN=1 python bench.py --steps 1 --warmup 3 --queries 64 --layout tex --knn-per-cell $PC --no-cpu --no-e2e --no-single 2>&1 | grep "ctp\] knn" | tail -1;  python bench.py --steps 5 --layout tex --knn-per-cell $PC --no-cpu --no-e2e --no-single | python tools/bsum.py pc$PC; done
