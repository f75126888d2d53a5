timeout 400 python -m pytest tests -m gpu -x -q -k "pairwise" -o faulthandler_timeout=60 2>&1 | tail -8
