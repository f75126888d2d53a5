python -m pytest tests -m gpu -x -q -k "tape or graph" 2>&1 | tail -2
python tools/probe_c5.py 2>&1 | tail -2
ADB_TAPE_GRAPH_STREAMS=1 python tools/probe_c5.py 2>&1 | tail -2
ADB_TAPE_GRAPH_STREAMS=8 python tools/probe_c5.py 2>&1 | tail -2
ADB_TAPE_SELFCHECK=1 ADB_TAPE_GRAPH_AFTER=0 ADB_TAPE_GRAPH_MIN_CALLS=1 python -m pytest tests -m gpu -x -q > gpurun_out/r3e_selfcheck.log 2>&1; echo selfcheck rc=$?; tail -3 gpurun_out/r3e_selfcheck.log
