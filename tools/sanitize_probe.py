"""Small end-to-end pass over the kernels changed this round, sized for compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rdis_b200 import Context, problems as P

# NonlinearProductFactor sweeps (value / gradient tile kernels, bulk store of the partials, gather)
spec = P.sinusoid(9, 2, 4)
ctx = Context.from_spec(spec); ctx.set_x(P.random_start(spec, 3))
s = ctx.eval()
g = ctx.grad()
print("sinusoid V=%d F=%d: sum %.12g |grad| %.12g" % (spec["V"], spec["F"], s, float(np.linalg.norm(g))))
ctx.close()
# bundle adjustment: point + camera block kernels, three visits (launch-order history), sweeps
ba = P.ba_synthetic(ncams=9, npts=700, nobs=4000, seed=5)
ctx = Context.from_spec(ba); x0 = ba["x0"]; ctx.set_x(x0)
pts, cams = P.ba_point_problems(ba), P.ba_camera_problems(ba)
bp, bc = ctx.batch(pts), ctx.batch(cams)
for visit in range(3):
    ctx.set_x(x0)
    bp.solve(x0[pts.vids].copy(), 25, 3e-8); rp = bp.fetch()
    bc.solve(None, 25, 3e-8); rc = bc.fetch()
print("ba: points %d (warps %d) sum %.12g, cameras %d sum %.12g, objective %.12g" % (pts.n, bp.info()["point_warps"], rp["f_end"].sum(), cams.n, rc["f_end"].sum(), ctx.eval()))
ctx.close()
