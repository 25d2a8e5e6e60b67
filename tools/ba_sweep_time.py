"""Times the BA all-factor sweeps (values; values + Jacobian rows) on 100 copies of the ladybug point cloud."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rdis_b200 import Context, problems as P
sp = P.ba_replicate_points(P.load_golden_ba(), 100)
ctx = Context.from_spec(sp); ctx.set_x(sp["x0"])
pf = torch.empty(sp["F"], dtype=torch.float64, device="cuda"); rows = torch.empty(sp["F"] * 12, dtype=torch.float64, device="cuda")
tot = torch.zeros(1, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream()
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps): fn()
    e.record(st); torch.cuda.synchronize(); return a.elapsed_time(e) / reps * 1e3
t1 = timed(lambda: ctx.eval_device(tot.data_ptr(), pf.data_ptr()))
t2 = timed(lambda: ctx.factor_rows_device(pf.data_ptr(), rows.data_ptr(), tot.data_ptr()))
b1, b2 = 32.0 * sp["F"] + 8.0 * sp["V"], 128.0 * sp["F"] + 8.0 * sp["V"]
print(os.environ.get("RDIS_B200_LIB", "default").split("/")[-1], "values %.1f us (%.3f of 6543.7 GB/s)  rows %.1f us (%.3f)  sum %.6f" % (t1, b1 / t1 / 1e3 / 6543.7, t2, b2 / t2 / 1e3 / 6543.7, float(tot.item())))
