python tools/probe_c5.py 2>&1 | tail -2
ADB_TAPE_GRAPH=0 python tools/probe_c5.py 2>&1 | tail -2
