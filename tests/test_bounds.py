"""Interval bounds: Factor::computeBounds / OptimizableFunction::computeBounds.

The reference holds NO golden bound values and Boost.Interval is not vendored (parity unpinned), so
the oracle's restatement is pinned by what the domain offers:
  - hand-derived intervals of small expressions (monotone pieces, sign cases of the product, the power quirk),
  - the ENCLOSURE property: for any point of the box, every factor's value lies inside its bound,
  - the all-assigned / assigned-constant rule of src/Factor.cpp:128 (the bound collapses to the point value).
The GPU test then compares rdisgpu_bounds with the oracle on the same inputs (1e-12: device sin/cos vs libm)."""
import numpy as np
import pytest


def _nlpf(rows, lb, ub):
    """rows: list of (coeff, [(vid, exponent, constant, sine), ...])"""
    rowptr = np.concatenate([[0], np.cumsum([len(r[1]) for r in rows])]).astype(np.int64)
    e = [t for r in rows for t in r[1]]
    return dict(kind="nlpf", V=len(lb), F=len(rows), lb=np.asarray(lb, float), ub=np.asarray(ub, float), rowptr=rowptr,
                vid=np.array([t[0] for t in e], np.int32), expo=np.array([t[1] for t in e], float),
                konst=np.array([t[2] for t in e], float), sine=np.array([t[3] for t in e], np.uint8),
                coeff=np.array([r[0] for r in rows], float))


HAND = _nlpf([
    (2.0, [(0, 1, 0, 0)]),                  # 2x, x in [-1, 3]                  -> [-2, 6]
    (1.0, [(1, 2, 0, 0)]),                  # y^2, y in [-2, 1]                 -> [0, 4]
    (1.0, [(1, 3, 0, 0)]),                  # y^3                               -> [-8, 1]
    (-1.0, [(0, 1, 0, 0), (1, 1, 0, 0)]),   # -(x*y): x*y in [-6, 3] (M * M)   -> [-3, 6]
    (1.0, [(2, 1, 0, 1)]),                  # sin z, z in [0.5, 1]              -> [sin .5, sin 1]
    (1.0, [(2, 1, -1.0, 1)]),               # sin(z + 1), z + 1 in [1.5, 2]: contains pi/2 -> [min(sin 1.5, sin 2), 1]
    (3.0, [(3, 1, 0, 1)]),                  # 3 sin w, w in [-10, 10]: a full period       -> [-3, 3]
    (1.0, [(4, -2, 0, 0)]),                 # u^-2, u in [1, 2]: the reference's power() returns 1/u -> [0.5, 1]
    (1.0, [(4, 4, 0.5, 0)]),                # (u - .5)^4 on [0.5, 1.5]          -> [.0625, 5.0625]
    (1.0, [(0, 2, 1.0, 0), (2, 1, 0, 0)]),  # (x-1)^2 * z: [0, 4] * [.5, 1]     -> [0, 4]
], lb=[-1, -2, 0.5, -10, 1], ub=[3, 1, 1.0, 10, 2])
HAND_WANT = [(-2, 6), (0, 4), (-8, 1), (-3, 6), (np.sin(0.5), np.sin(1.0)), (min(np.sin(1.5), np.sin(2.0)), 1.0), (-3, 3),
             (0.5, 1.0), (0.0625, 5.0625), (0, 4)]


def test_oracle_bounds_hand_derived(oracle_mod):
    orc = oracle_mod.OracleFunction.from_spec(HAND)
    lo, hi, tot = orc.bounds(np.zeros(HAND["V"], np.uint8))
    for k, (a, b) in enumerate(HAND_WANT):
        assert abs(lo[k] - a) <= 1e-12 and abs(hi[k] - b) <= 1e-12, (k, lo[k], hi[k], a, b)
    assert abs(tot[0] - lo.sum()) <= 1e-12 and abs(tot[1] - hi.sum()) <= 1e-12
    # src/Factor.cpp:128: every variable assigned -> the point value; an assigned constant -> the constant
    x = np.array([0.3, -1.2, 0.8, 2.5, 1.7])
    orc.set_x(x)
    _, pf = orc.eval(per_factor=True)
    lo, hi, _ = orc.bounds(np.ones(HAND["V"], np.uint8))
    assert np.array_equal(lo, pf) and np.array_equal(hi, pf)
    orc.set_factor_const(np.array([2]), np.array([7.5]), np.array([1], np.uint8))
    lo, hi, _ = orc.bounds(np.zeros(HAND["V"], np.uint8))
    assert lo[2] == 7.5 and hi[2] == 7.5


def _enclosure(orc, spec, point, x, rng, samples, tol=1e-9):
    """every sampled point of the box (assigned variables fixed at x) lands inside the factor bounds"""
    orc.set_x(x)
    lo, hi, tot = orc.bounds(point)
    assert not np.isnan(lo).any() and not np.isnan(hi).any()
    free = np.nonzero(point == 0)[0]
    worst = 0.0
    for _ in range(samples):
        y = x.copy()
        y[free] = rng.uniform(spec["lb"][free], spec["ub"][free])
        orc.set_x(y)
        s, pf = orc.eval(per_factor=True)
        slack = tol * np.maximum(1.0, np.abs(pf))
        assert (pf >= lo - slack).all() and (pf <= hi + slack).all(), int(np.argmax((pf < lo - slack) | (pf > hi + slack)))
        assert tot[0] - tol * max(1.0, abs(s)) <= s <= tot[1] + tol * max(1.0, abs(s))
        finite = np.isfinite(hi) & np.isfinite(lo)
        worst = max(worst, float(np.max((pf[finite] - lo[finite]) / np.maximum(hi[finite] - lo[finite], 1e-300))))
    orc.set_x(x)
    return lo, hi


def _random_nlpf(seed, V=30, F=120):
    rng = np.random.default_rng(seed)
    rows = []
    for _ in range(F):
        a = int(rng.integers(1, 5))
        vs = rng.choice(V, size=a, replace=False)
        rows.append((float(rng.normal(0, 2)), [(int(v), float(rng.choice([1, 1, 2, 3, 4])), float(rng.choice([0, 0, 0.7, -1.3])),
                                               int(rng.integers(0, 2))) for v in vs]))
    lb = rng.uniform(-4, 0, size=V); ub = lb + rng.uniform(0.01, 6, size=V)
    return _nlpf(rows, lb, ub)


def test_oracle_bounds_enclose_sampled_values(oracle_mod):
    from rdis_b200 import problems as P
    rng = np.random.default_rng(1)
    cases = [_random_nlpf(3), P.sinusoid(5, 2, 4)]
    ba = P.ba_synthetic(ncams=4, npts=30, nobs=100, seed=5)
    # boxes a branch & bound would meet late in the search: a few per cent around the state, depth bounded away from 0
    ba = dict(ba); w = 0.02 * np.maximum(np.abs(ba["x0"]), 0.05); ba["lb"] = ba["x0"] - w; ba["ub"] = ba["x0"] + w
    cases.append(ba)
    for sp in cases:
        orc = oracle_mod.OracleFunction.from_spec(sp)
        x = rng.uniform(sp["lb"], sp["ub"])
        for frac in (0.0, 0.5, 0.9):
            point = (rng.random(sp["V"]) < frac).astype(np.uint8)
            _enclosure(orc, sp, point, x, rng, samples=60)


def test_testpoly_bounds_contain_the_documented_optimum(oracle_mod):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "testpoly.npz"))
    spec = {k: g[k] for k in g.files}
    spec = dict(kind="nlpf", V=int(spec["V"]), F=int(spec["F"]), lb=spec["lb"], ub=spec["ub"], rowptr=spec["rowptr"], vid=spec["vid"],
                expo=spec["expo"], konst=spec["konst"], sine=spec["sine"], coeff=spec["coeff"])
    orc = oracle_mod.OracleFunction.from_spec(spec)
    lo, hi, tot = orc.bounds(np.zeros(spec["V"], np.uint8))
    assert tot[0] <= -168.2721 <= tot[1]      # data/testpoly.txt:18-22


@pytest.mark.gpu
def test_device_bounds_match_the_oracle(oracle_mod, built_lib):
    from rdis_b200 import Context, problems as P
    rng = np.random.default_rng(2)
    ba = P.ba_synthetic(ncams=4, npts=30, nobs=100, seed=5)
    tight = dict(ba); w = 0.02 * np.maximum(np.abs(ba["x0"]), 0.05); tight["lb"] = ba["x0"] - w; tight["ub"] = ba["x0"] + w
    cases = [HAND, _random_nlpf(3), _random_nlpf(4, V=200, F=3000), P.sinusoid(7, 2, 4), tight, ba]
    for sp in cases:
        ctx = Context.from_spec(sp); orc = oracle_mod.OracleFunction.from_spec(sp)
        x = rng.uniform(sp["lb"], sp["ub"])
        ctx.set_x(x); orc.set_x(x)
        for frac, fconst in ((0.0, ()), (0.5, (1, 3)), (0.9, ()), (1.0, ())):
            point = (rng.random(sp["V"]) < frac).astype(np.uint8) if frac < 1 else np.ones(sp["V"], np.uint8)
            if fconst:
                fid = np.array(fconst); val = np.array([0.25, -3.0]); on = np.ones(2, np.uint8)
                ctx.set_factor_const(fid, val, on); orc.set_factor_const(fid, val, on)
            lo, hi, tot = ctx.bounds(point)
            wlo, whi, wtot = orc.bounds(point)
            for got, want in ((lo, wlo), (hi, whi)):
                same_special = (np.isnan(got) == np.isnan(want)).all() and (np.isinf(got) == np.isinf(want)).all()
                assert same_special
                fin = np.isfinite(want)
                assert (np.sign(got[~fin & ~np.isnan(want)]) == np.sign(want[~fin & ~np.isnan(want)])).all()
                assert (np.abs(got[fin] - want[fin]) <= 1e-12 * np.maximum(1.0, np.abs(want[fin]))).all()
            for a, b in zip(tot, wtot):
                assert (np.isnan(a) and np.isnan(b)) or a == b or abs(a - b) <= 1e-11 * max(1.0, abs(b))
            # a sub-list, in the caller's order
            sub = rng.permutation(sp["F"])[: max(1, sp["F"] // 3)]
            slo, shi, _ = ctx.bounds(point, sub)
            assert np.array_equal(slo, lo[sub], equal_nan=True) and np.array_equal(shi, hi[sub], equal_nan=True)
            if fconst:
                off = np.zeros(2, np.uint8)
                ctx.set_factor_const(fid, val, off); orc.set_factor_const(fid, val, off)
