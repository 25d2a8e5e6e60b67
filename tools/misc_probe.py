"""Touches the kernels either side of the solve path once each, for ncu: interval bounds (per factor and per list), component
labelling, the cooperative-grid solve of one large component, the one-CTA LM kernel, the strict kernel.
usage: python tools/misc_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from rdis_b200 import Context, problems as P

spec = P.load_golden_ba(); x0 = spec["x0"]
pts, cams = P.ba_point_problems(spec), P.ba_camera_problems(spec)
ctx = Context.from_spec(spec); ctx.set_x(x0)
assigned = np.ones(spec["V"], np.uint8); assigned[9 * spec["ncams"]:] = 0          # points open
lo, hi, tot = ctx.bounds(assigned)
sums = ctx.bounds_lists(assigned, pts.fac_off, pts.fids)
vl, fl, ncomp, rounds = ctx.components(assigned)
print("bounds sum", tot, "lists", sums.shape, "components", ncomp, "rounds", rounds)
blk = bench.top_level_block(spec, P)
r = ctx.solve_cgd(blk, x0[blk.vids], 25, 3e-8)
print("grid solve f", r["f_init"][0], "->", r["f_end"][0], "evals", r["n_feval"][0])
ctx.set_x(x0)
r = ctx.solve_lm(cams, x0[cams.vids], 25, 3e-8)
print("LM cameras sum", r["f_end"].sum())
ctx.set_x(x0)
r = ctx.solve_lm(pts, x0[pts.vids], 25, 3e-8)
print("LM points sum", r["f_end"].sum())
sc = Context.from_spec(spec); sc.set_option("strict", 1); sc.set_x(x0)
r = sc.solve_cgd(pts, x0[pts.vids], 25, 3e-8)
r2 = sc.solve_cgd(cams, sc.get_x()[cams.vids], 25, 3e-8)
print("strict step objective", r2["f_end"].sum())
