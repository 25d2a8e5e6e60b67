"""One large top-level subspace solve (all variables of a ladybug-shaped graph / a 20 % block of them) through the
cooperative grid kernel, GPU vs the CPU oracle.  usage: python tools/toplevel_probe.py [--cpu]   (--cpu also runs the CPU oracle: minutes for the full problem)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rdis_b200 import Context, ProblemSet, problems as P  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

spec = P.ba_synthetic()
x0 = spec["x0"]
ctx = Context.from_spec(spec)
# (i) the whole problem; (ii) a top-level block: 10 cameras + 1500 points (~4.6k variables), factors = those fully inside
full = P.full_problem(spec)
cams = np.arange(10); pts = np.arange(1500)
vids = np.concatenate([(9 * cams[:, None] + np.arange(9)).ravel(), (9 * spec["ncams"] + 3 * pts[:, None] + np.arange(3)).ravel()]).astype(np.int32)
fids = np.nonzero(np.isin(spec["cam"], cams) | np.isin(spec["pt"], pts))[0].astype(np.int64)
block = ProblemSet([0, len(vids)], vids, [0, len(fids)], fids)
for name, ps in (("full problem", full), ("top-level block", block)):
    ctx.set_x(x0)
    b = ctx.batch(ps)
    for rep in range(2):
        ctx.set_x(x0); ctx.synchronize()
        t0 = time.perf_counter()
        b.solve(None, 25, 3e-8)
        ctx.synchronize()
        dt = time.perf_counter() - t0
    r = b.fetch()
    if "--cpu" in sys.argv:
        orc = O.OracleFunction.from_spec(spec); orc.set_x(x0)
        o = orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[ps.vids], 25, 3e-8)
    else:
        o = {"seconds": float("nan"), "f_end": np.array([float("nan")])}
    print("%s: nv %d nf %d | GPU %.2f ms (%d evaluations, %.1f us each) f %.6e -> %.9e | CPU oracle %.2f s f_end %.9e | rel diff %.2e | speed-up %.0fx | mapping %s" % (
        name, len(ps.vids), len(ps.fids), dt * 1e3, int(r["n_feval"][0]), dt * 1e6 / max(int(r["n_feval"][0]), 1), r["f_init"][0], r["f_end"][0],
        o["seconds"], o["f_end"][0], abs(r["f_end"][0] - o["f_end"][0]) / abs(o["f_end"][0]), o["seconds"] / dt, b.info()))
