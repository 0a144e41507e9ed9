timeout 200 python tools/lpc2_prof.py 2>&1 | tail -7
