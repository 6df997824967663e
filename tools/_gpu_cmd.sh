timeout 600 python -m pytest tests -m gpu -q -x -k "unidyn_slabs" 2>&1 | tail -15
