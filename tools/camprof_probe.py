import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from rdis_b200 import Context, problems as P
spec = P.load_golden_ba(); x0 = spec["x0"]; cams = P.ba_camera_problems(spec)
ctx = Context.from_spec(spec); ctx.set_x(x0)
r = ctx.solve_cgd(cams, x0[cams.vids], 25, 3e-8)
print("evals per camera: max", r["n_feval"].max(), "mean", r["n_feval"].mean(), "nf of the longest", np.diff(cams.fac_off)[np.argmax(r["n_feval"])])
