"""Camera-block kernel probe: evaluation counts, cluster-shape independence of the results, kernel time per shape."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from rdis_b200 import Context, problems as P
spec = P.load_golden_ba(); x0 = spec["x0"]; cams = P.ba_camera_problems(spec)
ref = None
for C in (0, 8, 7, 6, 5, 4, 2, 1):
    ctx = Context.from_spec(spec); ctx.set_option("camera_cluster", C); ctx.set_x(x0)
    b = ctx.batch(cams)
    ts = []
    for it in range(4):
        ctx.set_x(x0); ctx.synchronize()
        t0 = time.perf_counter(); b.solve(None, 25, 3e-8); ctx.synchronize(); ts.append(time.perf_counter() - t0)
    r = b.fetch()
    if ref is None:
        ref = r
        print("evals per camera: max", r["n_feval"].max(), "mean", r["n_feval"].mean(), "nf of the longest", np.diff(cams.fac_off)[np.argmax(r["n_feval"])])
    same = np.array_equal(r["f_end"].view(np.uint64), ref["f_end"].view(np.uint64)) and np.array_equal(r["x"].view(np.uint64), ref["x"].view(np.uint64))
    print("camera_cluster", C, b.info()["cluster_size"], "x", b.info()["camera_threads"], "ms %.3f" % (min(ts[1:]) * 1e3), "identical to the default shape:", same, "sum f_end %.9f" % r["f_end"].sum())
