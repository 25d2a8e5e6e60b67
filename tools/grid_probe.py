"""Runs the cooperative-grid solve of ladybug's 4755-variable top-level block once (for ncu captures / quick timing)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rdis_b200 import Context, problems as P
spec = P.load_golden_ba(); x0 = spec["x0"]
block = bench.top_level_block(spec, P)
ctx = Context.from_spec(spec); b = ctx.batch(block)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ctx.set_x(x0); ctx.synchronize()
    t0 = time.perf_counter(); b.solve(None, 25, 3e-8); ctx.synchronize(); t = time.perf_counter() - t0
r = b.fetch()
print("top-level block: %.3f ms, %d evaluations, %.2f us each, f_end %.9f" % (t * 1e3, r["n_feval"][0], t * 1e6 / r["n_feval"][0], r["f_end"][0]))
