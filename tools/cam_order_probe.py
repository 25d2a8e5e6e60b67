"""Camera-block kernel time against the ORDER of the camera problems in the batch (= cluster launch order): 49 clusters of
8 CTAs do not all fit at once (6 per GPC), so the last one starts when the first finishes."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from rdis_b200 import Context, problems as P
spec = P.load_golden_ba(); x0 = spec["x0"]; cams = P.ba_camera_problems(spec)
ctx = Context.from_spec(spec); ctx.set_x(x0)
def run(ps, tag):
    b = ctx.batch(ps); ts = []
    for it in range(5):
        ctx.set_x(x0); ctx.synchronize()
        t0 = time.perf_counter(); b.solve(None, 25, 3e-8); ctx.synchronize(); ts.append(time.perf_counter() - t0)
    r = b.fetch()
    print("%-28s %.3f ms  sum f_end %.9f" % (tag, min(ts[1:]) * 1e3, r["f_end"].sum()))
    return r
r = run(cams, "camera id order")
ev = r["n_feval"]; nf = np.diff(cams.fac_off)
print("evals:", ev.tolist()); print("nf   :", nf.tolist())
run(cams.subset(np.argsort(-ev, kind="stable")), "most evaluations first")
run(cams.subset(np.argsort(ev, kind="stable")), "fewest evaluations first")
run(cams.subset(np.argsort(-nf, kind="stable")), "most observations first")
run(cams.subset(np.argsort(-(ev * (nf + 2000.0)), kind="stable")), "predicted time first")
