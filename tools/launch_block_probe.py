"""Does rdisgpu_batch_solve_cgd return before the kernels finish?  (host time of the call vs time to synchronize)"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from rdis_b200 import Context, problems as P
spec = P.load_golden_ba(); x0 = spec["x0"]
ctx = Context.from_spec(spec); ctx.set_x(x0)
for name, ps in (("points", P.ba_point_problems(spec)), ("cameras", P.ba_camera_problems(spec))):
    b = ctx.batch(ps)
    xs = x0[ps.vids].copy()
    for mode in ("host x0", "no x0"):
        for it in range(4):
            ctx.set_x(x0); ctx.synchronize()
            t0 = time.perf_counter()
            b.solve(xs if mode == "host x0" else None, 25, 3e-8)
            t1 = time.perf_counter()
            ctx.synchronize()
            t2 = time.perf_counter()
        print("%-8s %-8s call %.3f ms, then synchronize %.3f ms" % (name, mode, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
