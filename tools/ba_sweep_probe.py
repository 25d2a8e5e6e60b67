"""BA all-factor sweeps (values; values + Jacobian rows) on 100 copies of the ladybug point cloud, for ncu.
usage: python tools/ba_sweep_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rdis_b200 import Context, problems as P  # noqa: E402

sp = P.ba_replicate_points(P.load_golden_ba(), 100)
ctx = Context.from_spec(sp)
ctx.set_x(sp["x0"])
pf = torch.empty(sp["F"], dtype=torch.float64, device="cuda")
rows = torch.empty(sp["F"] * 12, dtype=torch.float64, device="cuda")
tot = torch.zeros(1, dtype=torch.float64, device="cuda")
for _ in range(3):
    ctx.eval_device(tot.data_ptr(), pf.data_ptr())
    ctx.factor_rows_device(pf.data_ptr(), rows.data_ptr(), tot.data_ptr())
torch.cuda.synchronize()
print("sum", float(tot.item()))
