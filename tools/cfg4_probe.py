"""cfg4 (1e6 variables / 4e6 NonlinearProductFactors): sibling subtree solves after assigning the top tree levels.
usage: python tools/cfg4_probe.py [assigned_levels]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rdis_b200 import Context, problems as P  # noqa: E402

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 10
t0 = time.perf_counter()
spec = P.sinusoid(19, 2, 4)
x0 = P.random_start(spec, 1)
print("generated V=%d F=%d in %.1f s" % (spec["V"], spec["F"], time.perf_counter() - t0))
t0 = time.perf_counter()
ps = P.sinusoid_subtree_problems(spec, levels)
print("%d sibling components, %d vars / %d factors each, built in %.1f s" % (ps.n, ps.var_off[1], ps.fac_off[1], time.perf_counter() - t0))
t0 = time.perf_counter()
ctx = Context.from_spec(spec)
print("context (finalize incl. term table) in %.1f s" % (time.perf_counter() - t0))
ctx.set_x(x0)
f0 = ctx.eval()
results = {}
modes = os.environ.get("CFG4_MODES", "resident,generic").split(",")
reps = int(os.environ.get("CFG4_REPS", "2"))
for mode in modes:
    if mode == "generic":
        ctx.set_option("generic_only", 1)
    t0 = time.perf_counter()
    b = ctx.batch(ps)
    print(mode, "batch build %.1f ms, mapping" % ((time.perf_counter() - t0) * 1e3), b.info())
    for rep in range(reps):
        ctx.set_x(x0)
        ctx.synchronize()
        t0 = time.perf_counter()
        b.solve(None, 25, 3e-8)
        ctx.synchronize()
        dt = time.perf_counter() - t0
        r = b.fetch(want_x=True)
        print("%s rep %d: %.2f ms for %d solves => %.0f solves/s; objective %.6e -> %.6e; evals %d value + %d gradient" % (
            mode, rep, dt * 1e3, ps.n, ps.n / dt, f0, ctx.eval(), int(r["n_feval"].sum()), int(r["n_geval"].sum())))
    results[mode] = r
    b.close()
same = len(results) < 2 or all(np.array_equal(results["resident"][k], results["generic"][k]) for k in ("f_end", "iters", "status"))
print("resident == generic (f_end, iters, status):", same, "(equality is expected only with resident_threads=256; the default CTA is 512 wide)")
