#!/bin/bash
# times the sweep leg with every experiment build under gpurun_exp/ (kernel tuning only)
P='import json,sys; d=json.loads(sys.stdin.read()); print("%.2f us  frac %.3f   grad %.2f us frac %.3f" % (d["launch_ms"]*1e3, d["frac"], d["grad_sweep"]["launch_ms"]*1e3, d["grad_sweep"]["frac"]))'
echo -n "default: "; python tools/sweep_probe.py | python -c "$P"
for f in gpurun_exp/lib_*.so; do echo -n "$f: "; RDIS_B200_LIB=$PWD/$f python tools/sweep_probe.py 2>&1 | tail -1 | python -c "$P"; done
