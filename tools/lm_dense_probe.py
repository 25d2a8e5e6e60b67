"""LM on the 4755-variable top-level block of ladybug (for ncu: lm_syrk_kernel's tensor-pipe share).  usage: python tools/lm_dense_probe.py [itmax]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from rdis_b200 import Context, problems as P
spec = P.load_golden_ba(); x0 = spec["x0"]
blk = bench.top_level_block(spec, P)
ctx = Context.from_spec(spec); ctx.set_x(x0)
itmax = int(sys.argv[1]) if len(sys.argv) > 1 else 3
t0 = time.perf_counter(); r = ctx.solve_lm(blk, x0[blk.vids], itmax, 3e-8); t = time.perf_counter() - t0
print("m %d nf %d: %.1f ms, f %.6e -> %.6e, iters %d stop %d nfev %d" % (len(blk.vids), len(blk.fids), t * 1e3, r["f_init"][0], r["f_end"][0], r["iters"][0], r["stop"][0], r["n_feval"][0]))
