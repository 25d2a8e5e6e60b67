"""GPU leg of the parity attribution: the PRODUCTION kernels and the strict kernel on the real ladybug-49-7776 wave
against (a) the oracle's devtrig twin (same sin/cos as the device: isolates what the kernels themselves change) and
(b) the plain oracle (glibc sin/cos: what a user of the reference sees).  Run under gpurun:

    python tools/parity_gpu.py  ->  gpurun_out/r02_parity_gpu.json
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rdis_b200 import Context, problems as P
from oracle import oracle_py as O


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def cmp(r, o):
    rel = np.abs(r["f_end"] - o["f_end"]) / np.maximum(np.abs(o["f_end"]), 1e-300)
    return {"bit_identical": int((bits(r["f_end"]) == bits(o["f_end"])).sum()), "n": int(rel.size), "n_over_1e-6": int((rel > 1e-6).sum()),
            "max_rel": float(rel.max()), "median_rel": float(np.median(rel)), "iters_differ": int((r["iters"] != o["iters"]).sum()),
            "rel_of_sum": float(abs(r["f_end"].sum() - o["f_end"].sum()) / abs(o["f_end"].sum())), "sum_f_end": float(r["f_end"].sum())}


def main():
    spec = P.load_golden_ba()
    x0 = spec["x0"]
    out = {"graph": "data/ladybug-problem-49-7776-pre.txt, file state; SSmaxit 25, ftol 3e-8", "waves": {}}
    oracles = {}
    for variant in ("devtrig", "restated"):
        reps = [O.OracleFunction.from_spec(spec, variant) for _ in range(8)]
        for rp in reps:
            rp.set_x(x0)
        oracles[variant] = reps
    for name, ps in (("points", P.ba_point_problems(spec)), ("cameras", P.ba_camera_problems(spec))):
        row = {}
        ref = {}
        for variant, reps in oracles.items():
            for rp in reps:
                rp.set_x(x0)
            ref[variant] = reps[0].solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[ps.vids], 25, 3e-8, replicas=reps)
        row["oracle devtrig vs oracle glibc (libm alone)"] = cmp(ref["devtrig"], ref["restated"])
        for label, opts in (("production", {}), ("strict", {"strict": 1}), ("generic kernels", {"generic_only": 1})):
            ctx = Context.from_spec(spec)
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.set_x(x0)
            r = ctx.solve_cgd(ps, x0[ps.vids], 25, 3e-8)
            row[label + " vs oracle devtrig"] = cmp(r, ref["devtrig"])
            row[label + " vs oracle glibc"] = cmp(r, ref["restated"])
            row[label + " statuses"] = np.bincount(r["status"], minlength=8).tolist()
            row[label + " evals"] = int(r["n_feval"].sum())
        out["waves"][name] = row
        for k, v in row.items():
            print(name, "|", k, "|", v, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r02_parity_gpu.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
