python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -6
