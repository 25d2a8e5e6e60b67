"""Writes the golden fixtures under tests/golden/.  Run in the build container (needs
/root/reference for the data files and the oracle for parsing / reference-semantics results):

    python tests/golden/make_golden.py

ladybug_49_7776.npz   the reference's data/ladybug-problem-49-7776-pre.txt as parsed by the
                      oracle's restatement of BundleAdjustmentFunction::load
golden_solves.npz     oracle results (reference NR header build, oracle/_ref) for fixed subspace
                      problems; the GPU parity tests and the CPU oracle tests both check them
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data"


def main():
    b = O.OracleFunction.load_bal(os.path.join(REF, "ladybug-problem-49-7776-pre.txt"))
    sp = b.export()
    assert sp["ncams"] == 49 and sp["npts"] == 7776 and sp["F"] == 31843
    np.savez_compressed(os.path.join(HERE, "ladybug_49_7776.npz"), ncams=sp["ncams"], npts=sp["npts"],
                        cam=sp["cam"].astype(np.uint8), pt=sp["pt"].astype(np.uint16), obs=sp["obs"], x0=sp["x0"],
                        lb=sp["lb"], ub=sp["ub"])
    print("ladybug fixture:", os.path.getsize(os.path.join(HERE, "ladybug_49_7776.npz")), "bytes")

    # data/testpoly.txt as parsed by the oracle's restatement of PolynomialFunction::load
    tp = O.OracleFunction.load_poly(os.path.join(REF, "testpoly.txt")).export()
    np.savez_compressed(os.path.join(HERE, "testpoly.npz"), **{k: tp[k] for k in
                        ("V", "F", "lb", "ub", "rowptr", "vid", "expo", "konst", "sine", "coeff")})

    # golden subspace solves on the real ladybug graph, written from the build that drives the
    # solve with the reference's OWN minimize_nrc.h (oracle/_ref)
    from rdis_b200 import problems as P
    assert O.have_refnrc(), "golden solves must come from the reference-header build"
    spec = P.load_golden_ba()
    rng = np.random.default_rng(49_7776)
    sig = np.concatenate([np.tile([1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2, 0.2, 1e-9, 1e-15], 49), np.full(3 * 7776, 0.05)])
    x0 = np.clip(spec["x0"] + rng.normal(0, 1, spec["V"]) * sig, spec["lb"], spec["ub"])
    pts = P.ba_point_problems(spec).subset(range(0, 7776, 61))
    cams = P.ba_camera_problems(spec).subset([7, 30])
    out = {"x0": x0, "maxiters": 25, "ftol": 3e-8}
    for tag, ps in (("pts", pts), ("cams", cams)):
        orc = O.OracleFunction.from_spec(spec, "refnrc")
        orc.set_x(x0)
        o = orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[ps.vids], 25, 3e-8)
        out.update({tag + "_var_off": ps.var_off, tag + "_vids": ps.vids, tag + "_fac_off": ps.fac_off,
                    tag + "_fids": ps.fids, tag + "_f_init": o["f_init"], tag + "_f_end": o["f_end"],
                    tag + "_x": o["x"], tag + "_iters": o["iters"]})
        print(tag, ps.n, "solves; f_init sum %.6e -> f_end sum %.6e" % (o["f_init"].sum(), o["f_end"].sum()))
    orc = O.OracleFunction.from_spec(spec, "refnrc")
    orc.set_x(spec["x0"])
    out["f_file_x0"] = orc.eval()
    print("f(file x0) = %.10e" % out["f_file_x0"])
    np.savez_compressed(os.path.join(HERE, "golden_solves.npz"), **out)


if __name__ == "__main__":
    main()
