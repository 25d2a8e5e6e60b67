import sys, os
sys.path.insert(0, os.getcwd())
from rdis_b200 import Context, problems as P
spec = P.load_golden_ba(); x0 = spec["x0"]; cams = P.ba_camera_problems(spec)
ctx = Context.from_spec(spec); ctx.set_option("camera_cluster", 6); ctx.set_x(x0)
ctx.solve_cgd(cams, x0[cams.vids], 25, 3e-8)
