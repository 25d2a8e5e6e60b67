"""GPU probe: strict mode vs the devtrig oracle on the real ladybug wave (bit identity), sequence included."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from rdis_b200 import Context, problems as P
from oracle import oracle_py as O

def bits(a): return np.ascontiguousarray(a, np.float64).view(np.uint64)
def same(a, b): return int((bits(a) != bits(b)).sum())

spec = P.load_golden_ba(); x0 = spec["x0"]
pts, cams = P.ba_point_problems(spec), P.ba_camera_problems(spec)
ctx = Context.from_spec(spec); ctx.set_option("strict", 1); ctx.set_x(x0)
orc = O.OracleFunction.from_spec(spec, "devtrig"); orc.set_x(x0)
t = time.time(); r = ctx.solve_cgd(pts, x0[pts.vids], 25, 3e-8); tg = time.time() - t
o = orc.solve_cgd_batch(pts.var_off, pts.vids, pts.fac_off, pts.fids, x0[pts.vids], 25, 3e-8)
print("points: gpu %.1f ms cpu %.2f s; mismatching f_end %d f_init %d x %d iters %d (of %d)" % (
    tg * 1e3, o["seconds"], same(r["f_end"], o["f_end"]), same(r["f_init"], o["f_init"]), same(r["x"], o["x"]), int((r["iters"] != o["iters"]).sum()), pts.n))
bad = np.nonzero(bits(r["f_end"]) != bits(o["f_end"]))[0]
for k in bad[:8]:
    print("  comp", k, "nf", pts.fac_off[k+1]-pts.fac_off[k], "f_end gpu %.17g cpu %.17g iters %d/%d f_init same %s status %d nfe %d" % (
        r["f_end"][k], o["f_end"][k], r["iters"][k], o["iters"][k], bits(r["f_init"])[k] == bits(o["f_init"])[k], r["status"][k], r["n_feval"][k]))
# the camera wave AFTER the point wave on the same state (the caches persist on both sides)
xc = ctx.get_x(); 
print("state after points identical:", same(xc, orc.get_x()) == 0)
t = time.time(); r2 = ctx.solve_cgd(cams, xc[cams.vids], 25, 3e-8); tg = time.time() - t
o2 = orc.solve_cgd_batch(cams.var_off, cams.vids, cams.fac_off, cams.fids, xc[cams.vids], 25, 3e-8)
print("cameras after points: gpu %.1f ms cpu %.2f s; mismatching f_end %d f_init %d x %d iters %d (of %d)" % (
    tg * 1e3, o2["seconds"], same(r2["f_end"], o2["f_end"]), same(r2["f_init"], o2["f_init"]), same(r2["x"], o2["x"]), int((r2["iters"] != o2["iters"]).sum()), cams.n))
bad = np.nonzero(bits(r2["f_end"]) != bits(o2["f_end"]))[0]
for k in bad[:8]:
    print("  cam", k, "f_end gpu %.17g cpu %.17g iters %d/%d f_init same %s nfe %d" % (r2["f_end"][k], o2["f_end"][k], r2["iters"][k], o2["iters"][k], bits(r2["f_init"])[k] == bits(o2["f_init"])[k], r2["n_feval"][k]))
print("objective after the step: gpu %.17g cpu %.17g" % (r2["f_end"].sum(), o2["f_end"].sum()))
# NLPF: sibling subtrees + the cfg2 chain as one problem
for name, tree, lv in (("sinusoid h=9 subtrees", P.sinusoid(9, 2, 4), 3), ("cfg2 chain", P.sinusoid(999, 1, 3), 0)):
    xt = P.random_start(tree, 5)
    c2 = Context.from_spec(tree); c2.set_option("strict", 1); c2.set_x(xt)
    o2c = O.OracleFunction.from_spec(tree, "devtrig"); o2c.set_x(xt)
    ps = P.sinusoid_subtree_problems(tree, lv) if lv else P.full_problem(tree)
    t = time.time(); rr = c2.solve_cgd(ps, xt[ps.vids], 25, 3e-8); tg = time.time() - t
    oo = o2c.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, xt[ps.vids], 25, 3e-8)
    print("%s: %d problems gpu %.1f ms cpu %.2f s; mismatching f_end %d x %d iters %d; f_end[0] %.17g / %.17g" % (
        name, ps.n, tg * 1e3, oo["seconds"], same(rr["f_end"], oo["f_end"]), same(rr["x"], oo["x"]), int((rr["iters"] != oo["iters"]).sum()), rr["f_end"][0], oo["f_end"][0]))
