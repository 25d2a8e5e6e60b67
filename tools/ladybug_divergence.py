import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from rdis_b200 import Context, problems as P
from rdis_b200.capi import DONE_NAMES
from oracle import oracle_py as O
spec = P.load_golden_ba(); x0 = spec["x0"]; pts = P.ba_point_problems(spec)
ctx = Context.from_spec(spec); ctx.set_x(x0)
r = ctx.solve_cgd(pts, x0[pts.vids], 25, 3e-8)
g = Context.from_spec(spec); g.set_option("generic_only", 1); g.set_x(x0)
rg = g.solve_cgd(pts, x0[pts.vids], 25, 3e-8)
orc = O.OracleFunction.from_spec(spec); orc.set_x(x0)
o = orc.solve_cgd_batch(pts.var_off, pts.vids, pts.fac_off, pts.fids, x0[pts.vids], 25, 3e-8)
ot = O.OracleFunction.from_spec(spec, "fma"); ot.set_x(x0)
t = ot.solve_cgd_batch(pts.var_off, pts.vids, pts.fac_off, pts.fids, x0[pts.vids], 25, 3e-8)
rel = np.abs(r["f_end"] - o["f_end"]) / np.maximum(np.abs(o["f_end"]), 1e-12)
relg = np.abs(rg["f_end"] - o["f_end"]) / np.maximum(np.abs(o["f_end"]), 1e-12)
relt = np.abs(t["f_end"] - o["f_end"]) / np.maximum(np.abs(o["f_end"]), 1e-12)
print("fast within1e-6", (rel <= 1e-6).sum(), "generic", (relg <= 1e-6).sum(), "twin", (relt <= 1e-6).sum())
nf = np.diff(pts.fac_off)
for k in np.argsort(-rel)[:14]:
    print("comp %5d nf %2d  f_init %.6e | f_end gpu %.9e generic %.9e cpu %.9e twin %.9e | iters gpu %d gen %d cpu %d twin %d | status %s | rel %.2e relgen %.2e reltwin %.2e | x clipped %s" % (
        k, nf[k], r["f_init"][k], r["f_end"][k], rg["f_end"][k], o["f_end"][k], t["f_end"][k], r["iters"][k], rg["iters"][k], o["iters"][k], t["iters"][k],
        DONE_NAMES[r["status"][k]], rel[k], relg[k], relt[k],
        bool(np.any((r["x"][3*k:3*k+3] <= spec["lb"][pts.vids[3*k:3*k+3]]) | (r["x"][3*k:3*k+3] >= spec["ub"][pts.vids[3*k:3*k+3]])))))
print("f_init rel max", (np.abs(r["f_init"] - o["f_init"]) / np.abs(o["f_init"])).max())
print("gpu better than cpu on", (r["f_end"] < o["f_end"]).sum(), "worse on", (r["f_end"] > o["f_end"]).sum())
bad = rel > 1e-6
print("among bad: gpu lower objective on", (r["f_end"][bad] < o["f_end"][bad]).sum(), "of", bad.sum())
