python -m pytest tests -m gpu -q -x --timeout 300 -k "sample_multiple" 2>&1 | tail -3
for PCH in 16 32 64 128; do python bench.py --steps 5 --layout tex --no-cpu --no-single --pipe-chunk $PCH | python tools/bsum.py chunk$PCH; done
