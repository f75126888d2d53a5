timeout 300 python -m pytest tests -m gpu -x -q -k "conv" -o faulthandler_timeout=120 2>&1 | tail -15
timeout 200 python tools/bench_c4.py 2>&1 | tail -3
ADB_IMPLICIT_CONV=0 timeout 200 python tools/bench_c4.py 2>&1 | tail -3
