"""Interval bounds: Factor::computeBounds / OptimizableFunction::computeBounds.

The reference holds NO golden bound values and Boost.Interval is not vendored (parity unpinned), so
the oracle's restatement is pinned by what the domain offers:
  - hand-derived intervals of small expressions (monotone pieces, sign cases of the product, the power quirk),
  - the ENCLOSURE property: for any point of the box, every factor's value lies inside its bound — which is the
    reference's own DEBUG invariant for an optimised child component (src/Component.cpp:259-305, tol 1e-6),
  - an independent interval library (mpmath.iv) evaluating the same expressions,
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


def _iv_close(lo, hi, want, tol=1e-11):
    a, b = float(want.a), float(want.b)
    return abs(lo - a) <= tol * max(1.0, abs(a)) and abs(hi - b) <= tol * max(1.0, abs(b))


def test_oracle_bounds_match_an_independent_interval_library(oracle_mod):
    """Cross-check of the Boost.Interval restatement against mpmath's interval context (`mpmath.iv`: outward-rounded,
    30 digits), an independent implementation: the natural interval extension of the same expressions in the same
    order.  Every primitive the bounds use is range-tight in both libraries (products, integer powers, square with a
    zero-straddling argument, sqrt, division by a zero-free interval, sin / cos on arbitrary widths), so the two agree
    to rounding.  Covers random NonlinearProductFactors (mixed-sign domains, exponents 1..4, constants, sine flags) and
    bundle-adjustment factors on boxes a few per cent wide (the expression order of
    BundleAdjustmentFactor.cpp:104-157,186-232 written out again here)."""
    mpmath = pytest.importorskip("mpmath")
    from mpmath import iv
    from rdis_b200 import problems as P
    iv.dps = 30
    rng = np.random.default_rng(11)
    # ---- NonlinearProductFactor ----
    sp = _random_nlpf(21, V=25, F=150)
    orc = oracle_mod.OracleFunction.from_spec(sp)
    x = rng.uniform(sp["lb"], sp["ub"])
    orc.set_x(x)
    point = (rng.random(sp["V"]) < 0.3).astype(np.uint8)
    lo, hi, _ = orc.bounds(point)
    for f in range(sp["F"]):
        if all(point[sp["vid"][e]] for e in range(sp["rowptr"][f], sp["rowptr"][f + 1])):
            continue  # all assigned: the point value (checked elsewhere)
        acc = iv.mpf(1)
        for e in range(sp["rowptr"][f], sp["rowptr"][f + 1]):
            v = sp["vid"][e]
            val = iv.mpf(float(x[v])) if point[v] else iv.mpf([float(sp["lb"][v]), float(sp["ub"][v])])
            if sp["konst"][e] != 0:
                val = val - iv.mpf(float(sp["konst"][e]))
            if sp["expo"][e] != 1:
                val = val ** int(sp["expo"][e])
            if sp["sine"][e]:
                val = iv.sin(val)
            acc = acc * val
        acc = acc * iv.mpf(float(sp["coeff"][f]))
        assert _iv_close(lo[f], hi[f], acc), (f, lo[f], hi[f], acc)
    # ---- bundle adjustment ----
    ba = P.ba_synthetic(ncams=3, npts=20, nobs=50, seed=9)
    ba = dict(ba); w = 0.02 * np.maximum(np.abs(ba["x0"]), 0.05); ba["lb"] = ba["x0"] - w; ba["ub"] = ba["x0"] + w
    orc = oracle_mod.OracleFunction.from_spec(ba)
    orc.set_x(ba["x0"])
    point = (rng.random(ba["V"]) < 0.4).astype(np.uint8)
    lo, hi, _ = orc.bounds(point)
    nc = ba["ncams"]

    def var(v):
        return iv.mpf(float(ba["x0"][v])) if point[v] else iv.mpf([float(ba["lb"][v]), float(ba["ub"][v])])

    checked = 0
    for f in range(ba["F"]):
        c, p = int(ba["cam"][f]), int(ba["pt"][f])
        vids = [9 * c + s for s in range(9)] + [9 * nc + 3 * p + d for d in range(3)]
        if all(point[v] for v in vids):
            continue
        vals = [var(v) for v in vids]
        pt = vals[9:12]
        theta = iv.sqrt(vals[0] ** 2 + vals[1] ** 2 + vals[2] ** 2)
        vv = [vals[i] / theta for i in range(3)]
        if float(theta.b) - float(theta.a) < 1e-6:
            m = (float(theta.a) + float(theta.b)) / 2.0
            ct, st = iv.mpf(float(np.cos(m))), iv.mpf(float(np.sin(m)))
        else:
            ct, st = iv.cos(theta), iv.sin(theta)
        om = iv.mpf(1) - ct
        vxp = [vv[1] * pt[2] - vv[2] * pt[1], vv[2] * pt[0] - vv[0] * pt[2], vv[0] * pt[1] - vv[1] * pt[0]]
        vdp = vv[0] * pt[0] + vv[1] * pt[1] + vv[2] * pt[2]
        q = [pt[i] * ct + vxp[i] * st + vv[i] * om * vdp for i in range(3)]
        q = [q[0] + vals[3], q[1] + vals[4], q[2] + vals[5]]
        if float(q[2].a) <= 0.0 <= float(q[2].b):
            continue  # depth straddles zero: half-lines / whole line, library conventions differ
        px, py = -q[0] / q[2], -q[1] / q[2]
        r2 = px ** 2 + py ** 2
        dstn = iv.mpf(1) + vals[7] * r2 + vals[8] * r2 ** 2
        px, py = vals[6] * dstn * px, vals[6] * dstn * py
        err = ((px - iv.mpf(float(ba["obs"][f][0]))) ** 2 + (py - iv.mpf(float(ba["obs"][f][1]))) ** 2) / 2
        assert _iv_close(lo[f], hi[f], err, 1e-9), (f, lo[f], hi[f], err)
        checked += 1
    assert checked >= 20


def test_component_optimum_lies_inside_its_unassigned_bounds(oracle_mod):
    """The reference's own DEBUG invariant (Component::onChildEvaluated, src/Component.cpp:259-305, tol 1e-6): the value a
    child component is optimised to lies inside the bounds computed for it while its variables were unassigned.
    Sibling subtrees of the sinusoid family (ancestors assigned = points, the subtree's variables at their domains),
    optimised by the CGD oracle; and bundle-adjustment point blocks on boxes a few per cent wide."""
    from rdis_b200 import problems as P
    tol = 1e-6
    tree = P.sinusoid(7, 2, 4)
    x0 = P.random_start(tree, 13)
    ps = P.sinusoid_subtree_problems(tree, 3)
    orc = oracle_mod.OracleFunction.from_spec(tree); orc.set_x(x0)
    point = np.ones(tree["V"], np.uint8); point[ps.vids] = 0
    bounds = []
    for k in range(ps.n):
        fids = ps.fids[ps.fac_off[k]:ps.fac_off[k + 1]]
        bounds.append(orc.bounds(point, fids)[2])
    o = orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[ps.vids], 25, 3e-8)
    for k, (lo, hi) in enumerate(bounds):
        assert lo - tol <= o["f_end"][k] <= hi + tol and lo - tol <= o["f_init"][k] <= hi + tol, (k, lo, o["f_end"][k], hi)
    ba = P.ba_synthetic(ncams=4, npts=40, nobs=140, seed=6)
    ba = dict(ba); w = 0.03 * np.maximum(np.abs(ba["x0"]), 0.05); ba["lb"] = ba["x0"] - w; ba["ub"] = ba["x0"] + w
    pts = P.ba_point_problems(ba)
    orc = oracle_mod.OracleFunction.from_spec(ba); orc.set_x(ba["x0"])
    point = np.ones(ba["V"], np.uint8); point[pts.vids] = 0
    bounds = [orc.bounds(point, pts.fids[pts.fac_off[k]:pts.fac_off[k + 1]])[2] for k in range(pts.n)]
    o = orc.solve_cgd_batch(pts.var_off, pts.vids, pts.fac_off, pts.fids, ba["x0"][pts.vids], 25, 3e-8)
    for k, (lo, hi) in enumerate(bounds):
        assert lo - tol * max(1.0, abs(lo)) <= o["f_end"][k] <= hi + tol * max(1.0, abs(hi)), (k, lo, o["f_end"][k], hi)


@pytest.mark.gpu
def test_list_bounds_on_device_match_the_per_list_calls(built_lib, oracle_mod):
    """rdisgpu_bounds_lists (the unassigned bounds of EVERY child of a decomposition in one launch, each list's interval
    sum folded on the device in list order) against one rdisgpu_bounds call per list (bit for bit: same per-factor
    arithmetic, same fold order) and against the oracle (1e-12): sinusoid subtrees and bundle-adjustment point blocks,
    lists shorter and longer than a warp, an empty list."""
    from rdis_b200 import Context, problems as P
    tree = P.sinusoid(8, 2, 4)
    x0 = P.random_start(tree, 3)
    ps = P.sinusoid_subtree_problems(tree, 3)
    ctx = Context.from_spec(tree); ctx.set_x(x0)
    orc = oracle_mod.OracleFunction.from_spec(tree); orc.set_x(x0)
    assigned = np.ones(tree["V"], np.uint8); assigned[ps.vids] = 0
    off = np.concatenate([ps.fac_off, [ps.fac_off[-1]]])            # + one empty list at the end
    sums = ctx.bounds_lists(assigned, off, ps.fids)
    assert sums.shape == (ps.n + 1, 2) and sums[-1, 0] == 0.0 and sums[-1, 1] == 0.0
    for k in range(ps.n):
        fids = ps.fids[ps.fac_off[k]:ps.fac_off[k + 1]]
        _, _, tot = ctx.bounds(assigned, fids)
        assert sums[k, 0] == tot[0] and sums[k, 1] == tot[1]
        wlo, whi = orc.bounds(assigned, fids)[2]
        assert abs(sums[k, 0] - wlo) <= 1e-12 * max(1.0, abs(wlo)) and abs(sums[k, 1] - whi) <= 1e-12 * max(1.0, abs(whi))
    ba = P.ba_synthetic(ncams=4, npts=60, nobs=230, seed=6)
    ba = dict(ba); w = 0.03 * np.maximum(np.abs(ba["x0"]), 0.05); ba["lb"] = ba["x0"] - w; ba["ub"] = ba["x0"] + w
    pts = P.ba_point_problems(ba)
    bctx = Context.from_spec(ba); bctx.set_x(ba["x0"])
    assigned = np.ones(ba["V"], np.uint8); assigned[pts.vids] = 0
    sums = bctx.bounds_lists(assigned, pts.fac_off, pts.fids)
    for k in range(0, pts.n, 7):
        _, _, tot = bctx.bounds(assigned, pts.fids[pts.fac_off[k]:pts.fac_off[k + 1]])
        assert sums[k, 0] == tot[0] and sums[k, 1] == tot[1]
