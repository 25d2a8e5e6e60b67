"""Evaluation-count statistics of the bench step's two sibling batches (how long the serial chains are).
usage: python tools/solve_stats.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from rdis_b200 import Context, problems as P  # noqa: E402

spec = P.ba_synthetic(seed=bench.SEED)
x0 = spec["x0"]
pts, cams = P.ba_point_problems(spec), P.ba_camera_problems(spec)
ctx = Context.from_spec(spec)
for name, ps in (("points", pts), ("cameras", cams)):
    ctx.set_x(x0)
    b = ctx.batch(ps)
    for _ in range(3):
        ctx.set_x(x0)
        b.solve(None, 25, 3e-8)
    ctx.synchronize()
    ts = []
    for _ in range(5):
        ctx.set_x(x0)
        ctx.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        b.solve(None, 25, 3e-8)
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    r = b.fetch()
    nf, ng = r["n_feval"], r["n_geval"]
    ms = float(np.median(ts))
    print("%s: kernel %.3f ms; evals/problem mean %.1f max %d (with slope: mean %.1f max %d); iters mean %.1f max %d; "
          "=> %.2f us per evaluation on the longest chain; status histogram %s" %
          (name, ms, nf.mean(), nf.max(), ng.mean(), ng.max(), r["iters"].mean(), r["iters"].max(), ms * 1e3 / nf.max(),
           np.bincount(r["status"], minlength=8).tolist()))
