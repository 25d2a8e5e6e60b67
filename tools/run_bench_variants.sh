#!/bin/bash
# main bench (solve step) with every experiment build under gpurun_exp/ (kernel tuning only)
P='import json,sys; d=json.loads(sys.stdin.read()); print("%.0f solves/s  step %.3f ms  kernels %s  cams %s obj %.9e" % (d["value"], d["ms_per_step"], {k: round(v,3) for k,v in d["kernel_ms"].items()}, {k: d["config"]["mapping"]["cameras"][k] for k in ("cluster_size","camera_threads")}, d["objective_after_step"]))'
echo -n "default: "; python bench.py --steps 10 --warmup 3 --no-cpu --no-sweep | tail -1 | python -c "$P"
for f in gpurun_exp/lib_*.so; do echo -n "$f: "; RDIS_B200_LIB=$PWD/$f python bench.py --steps 10 --warmup 3 --no-cpu --no-sweep 2>&1 | tail -1 | python -c "$P"; done
