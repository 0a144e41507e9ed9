for nh in 2 3 1; do echo "== NH=$nh"; QCQP_LPC2_NH=$nh timeout 300 python tools/lpc2_probe.py 1000 1024 3 2>&1 | grep -v "L2 ->\|FP64"; done
QCQP_LPC2_NH=2 timeout 120 python tools/lpc2_probe.py 333 200 1 2>&1 | grep "bit-ident"
QCQP_LPC2_NH=3 timeout 120 python tools/lpc2_probe.py 130 64 1 2>&1 | grep "bit-ident"
QCQP_LPC2_NH=2 timeout 120 python tools/lpc2_probe.py 2100 64 1 2>&1 | grep "bit-ident\|LPC2=1"
