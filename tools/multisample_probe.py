"""Throughput headroom when several independent ladybug-shaped problems (optBA's --nsamples restarts) share one batch:
K copies of the graph in one context, one point wave + one camera wave over all copies.
usage: python tools/multisample_probe.py K [K ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from rdis_b200 import Context, problems as P  # noqa: E402


base = P.ba_synthetic(seed=bench.SEED)
for K in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    spec = P.ba_replicate(base, K)
    x0 = spec["x0"]
    pts, cams = P.ba_point_problems(spec), P.ba_camera_problems(spec)
    ctx = Context.from_spec(spec)
    bp, bc = ctx.batch(pts), ctx.batch(cams)
    x0d = torch.from_numpy(x0).cuda()
    ts = []
    for rep in range(6):
        ctx.set_x_device(x0d.data_ptr(), spec["V"])
        torch.cuda.synchronize()
        a, m, e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        bp.solve(None, 25, 3e-8)
        m.record()
        bc.solve(None, 25, 3e-8)
        e.record()
        torch.cuda.synchronize()
        ts.append((a.elapsed_time(m), m.elapsed_time(e)))
    tp, tc = np.median([t[0] for t in ts[2:]]), np.median([t[1] for t in ts[2:]])
    n = pts.n + cams.n
    print("K=%d: points %.3f ms cameras %.3f ms => %.0f solves/s (%.2fx of K=1 per-problem time); camera mapping %s" % (
        K, tp, tc, n / ((tp + tc) * 1e-3), (tp + tc), {k: bc.info()[k] for k in ("cluster_size", "camera_threads")}))
