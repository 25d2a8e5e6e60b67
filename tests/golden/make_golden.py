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


if __name__ == "__main__":
    main()
