"""Point-block kernel probe for the RDIS_PT_PROFILE build (device printf of per-pass cycle counts of the longest warps).
usage: RDIS_B200_LIB=build_exp/libptprof.so python tools/ptprof_probe.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from rdis_b200 import Context, problems as P
spec = P.load_golden_ba(); x0 = spec["x0"]; pts = P.ba_point_problems(spec)
ctx = Context.from_spec(spec); ctx.set_x(x0)
b = ctx.batch(pts)
b.solve(x0[pts.vids].copy(), 25, 3e-8); ctx.synchronize()
r = b.fetch()
nf = np.diff(pts.fac_off)
top = np.argsort(-r["n_feval"])[:8]
print("longest point blocks: evals", r["n_feval"][top], "observations", nf[top])
