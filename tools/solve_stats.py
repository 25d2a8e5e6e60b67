"""Diagnostic (GPU box): per-class statistics of the bench workload's solves."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rdis_b200 import Context, problems as P

spec = P.ba_synthetic(seed=20260417)
x0 = spec["x0"]
ctx = Context.from_spec(spec)
for name, ps in (("points", P.ba_point_problems(spec)), ("cameras", P.ba_camera_problems(spec))):
    ctx.set_x(x0)
    b = ctx.batch(ps)
    for _ in range(2):
        ctx.set_x(x0); b.solve(None, 25, 3e-8); ctx.synchronize()
    ctx.set_x(x0)
    t0 = time.perf_counter(); b.solve(None, 25, 3e-8); ctx.synchronize(); dt = time.perf_counter() - t0
    r = b.fetch()
    nf = np.diff(ps.fac_off)
    print(name, "n", ps.n, "ms %.3f" % (dt * 1e3), "nf min/med/max", nf.min(), np.median(nf), nf.max())
    print("  n_feval mean %.1f max %d  n_geval mean %.1f  iters mean %.1f max %d" % (
        r["n_feval"].mean(), r["n_feval"].max(), r["n_geval"].mean(), r["iters"].mean(), r["iters"].max()))
    print("  status hist", np.bincount(r["status"], minlength=8))
    print("  factor-evals total %.3e ; per ms %.3e" % ((r["n_feval"] * nf).sum(), (r["n_feval"] * nf).sum() / (dt * 1e3)))
    if name == "cameras":
        print("  per-problem: nf, n_feval", list(zip(nf.tolist(), r["n_feval"].tolist()))[:12])
