"""Attribution of solve differences on the REAL ladybug-49-7776 wave, one switch at a time (CPU only).

Every row runs the full point wave (7776 solves) and camera wave (49 solves) of the reference's ladybug file from
the file's state with one perturbation twin of the oracle (oracle/Makefile) and compares final objectives with the
plain oracle (no FMA, glibc sin/cos, true divisions, left-to-right sums, change filter on).  It answers: how far
does the REFERENCE ITSELF move under each of the benign changes a GPU implementation might make?

    python tools/parity_attribution.py [--threads N]  ->  profiles/r02_parity_attribution.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
import importlib.util
_ps = importlib.util.spec_from_file_location("rdis_problems", os.path.join(ROOT, "rdis_b200", "problems.py"))
P = importlib.util.module_from_spec(_ps)
_ps.loader.exec_module(P)  # numpy only: librdis_b200.so is not mapped


def wave(variant, spec, ps, x0, filt=True):
    O.set_change_filter(filt, variant)
    orc = O.OracleFunction.from_spec(spec, variant)
    orc.set_x(x0)
    r = orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[ps.vids], 25, 3e-8)
    O.set_change_filter(True, variant)
    return r


def compare(r, base):
    rel = np.abs(r["f_end"] - base["f_end"]) / np.maximum(np.abs(base["f_end"]), 1e-300)
    return {"n_over_1e-6": int((rel > 1e-6).sum()), "max_rel": float(rel.max()), "median_rel": float(np.median(rel)),
            "sum_f_end": float(r["f_end"].sum()), "rel_of_sum": float(abs(r["f_end"].sum() - base["f_end"].sum()) / base["f_end"].sum()),
            "iters_differ": int((r["iters"] != base["iters"]).sum())}


def main():
    spec = P.load_golden_ba()
    x0 = spec["x0"]
    out = {"what": __doc__.split("\n")[0], "graph": "data/ladybug-problem-49-7776-pre.txt (tests/golden/ladybug_49_7776.npz)", "rows": {}}
    for name, ps in (("points", P.ba_point_problems(spec)), ("cameras", P.ba_camera_problems(spec))):
        t0 = time.time()
        base = wave("restated", spec, ps, x0)
        rows = {"_baseline": {"n": int(ps.n), "sum_f_end": float(base["f_end"].sum()), "cpu_seconds": time.time() - t0}}
        rows["change filter off (src/Variable.cpp:69-73 ignored)"] = compare(wave("restated", spec, ps, x0, filt=False), base)
        for variant, label in (("fma", "FMA contraction on (gcc -mfma -ffp-contract=fast)"),
                               ("recip", "24 gradient quotients as reciprocal products"),
                               ("treefold", "factor values summed as a balanced tree"),
                               ("devtrig", "sin/cos = CUDA's algorithm instead of glibc's"),
                               ("devtrig_treefold", "CUDA sin/cos + tree sum")):
            rows[label] = compare(wave(variant, spec, ps, x0), base)
            print(name, label, rows[label], flush=True)
        out["rows"][name] = rows
    path = os.path.join(ROOT, "profiles", "r02_parity_attribution.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
