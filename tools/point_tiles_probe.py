"""Point-block kernel time against the number of point blocks that share a warp (divergence of the state-machine steps)."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from rdis_b200 import Context, problems as P
spec = P.load_golden_ba(); x0 = spec["x0"]; pts = P.ba_point_problems(spec)
frac = int(sys.argv[1]) if len(sys.argv) > 1 else 1  # every frac-th point block only (what one of `frac` ranks holds)
if frac > 1:
    pts = pts.subset(np.arange(0, pts.n, frac))
    print("subset: every %d-th point block, %d problems" % (frac, pts.n))
ref = None
for cap in (0, 32, 8, 4, 2, 1):
    ctx = Context.from_spec(spec); ctx.set_option("point_tiles_per_warp", cap); ctx.set_x(x0)
    b = ctx.batch(pts); ts = []
    for it in range(5):
        ctx.set_x(x0); ctx.synchronize()
        t0 = time.perf_counter(); b.solve(None, 25, 3e-8); ctx.synchronize(); ts.append(time.perf_counter() - t0)
    r = b.fetch()
    if ref is None: ref = r
    print("tiles per warp <= %2d: warps %5d  %.3f ms  identical results: %s" % (cap, b.info()["point_warps"], min(ts[1:]) * 1e3,
          np.array_equal(r["f_end"].view(np.uint64), ref["f_end"].view(np.uint64)) and np.array_equal(r["x"].view(np.uint64), ref["x"].view(np.uint64))))
