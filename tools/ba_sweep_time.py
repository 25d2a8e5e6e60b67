import os,sys
sys.path.insert(0, os.getcwd())
import torch, bench
from rdis_b200 import Context, problems as P
sp = P.ba_replicate_points(P.ba_synthetic(seed=bench.SEED), 100)
ctx = Context.from_spec(sp); ctx.set_x(sp["x0"])
pf = torch.empty(sp["F"], dtype=torch.float64, device="cuda"); tot = torch.zeros(1, dtype=torch.float64, device="cuda")
for _ in range(5): ctx.eval_device(tot.data_ptr(), pf.data_ptr())
torch.cuda.synchronize()
a,e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): ctx.eval_device(tot.data_ptr(), pf.data_ptr())
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e)/20
print("%.2f us  %.3f of peak" % (ms*1e3, (32.0*sp["F"]+8.0*sp["V"])/(ms*1e-3)/1e9/6543.7))
