"""GPU parity tests proper: every call goes through the C-ABI (rdis_b200.capi -> librdis_b200.so)
and is compared with the CPU oracle on the same seeded inputs.

Tolerances (written here as the contract):
  per-factor value / partial derivative   1e-12 relative (+1e-12 * scale absolute for cancellation)
  sums over factors                       1e-12 relative
  final objective of a subspace solve     1e-6 relative (north_star), typically observed ~1e-12
  index bookkeeping (which variables moved, statuses of empty problems)   bit-exact
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(built_lib):
    from rdis_b200 import Context  # noqa: F401
    import rdis_b200
    return rdis_b200


def _relerr(a, b, floor=1e-300):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def _ba_small(P):
    return P.ba_synthetic(ncams=7, npts=120, nobs=520, seed=11)


def test_eval_and_grad_ba(gpu, oracle_mod):
    from rdis_b200 import Context, problems as P
    spec = _ba_small(P)
    ctx = Context.from_spec(spec); orc = oracle_mod.OracleFunction.from_spec(spec)
    ctx.set_x(spec["x0"]); orc.set_x(spec["x0"])
    sg, pg = ctx.eval(per_factor=True)
    so, po = orc.eval(per_factor=True)
    assert _relerr(pg, po).max() <= 1e-12
    assert abs(sg - so) <= 1e-12 * abs(so)
    # subset in caller order
    fid = np.random.default_rng(0).permutation(spec["F"])[:97]
    assert abs(ctx.eval(fid) - orc.eval(fid)) <= 1e-12 * abs(orc.eval(fid))
    # per-factor Jacobian rows
    rows = ctx.factor_grad(np.arange(spec["F"]), 12)
    ref = np.stack([orc.factor_grad(j, 12) for j in range(spec["F"])])
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(rows - ref) <= 1e-11 * scale + 1e-300).all()
    # accumulated gradient, all factors / all vars and subset / subset
    g = ctx.grad(); go = orc.grad()
    assert (np.abs(g - go) <= 1e-11 * np.abs(go).max()).all()
    vid = np.arange(9 * 3, 9 * 3 + 9)
    fsub = np.nonzero(spec["cam"] == 3)[0][::2]
    g2 = ctx.grad(fsub, vid); go2 = orc.grad(fsub, vid)
    assert (np.abs(g2 - go2) <= 1e-11 * np.abs(go2).max()).all()


def test_eval_and_grad_nlpf(gpu, oracle_mod):
    from rdis_b200 import Context, problems as P
    spec = P.sinusoid(6, 3, 4, odd=True)
    x0 = P.random_start(spec, 3)
    ctx = Context.from_spec(spec); orc = oracle_mod.OracleFunction.from_spec(spec)
    ctx.set_x(x0); orc.set_x(x0)
    sg, pg = ctx.eval(per_factor=True)
    so, po = orc.eval(per_factor=True)
    assert (np.abs(pg - po) <= 1e-13 * np.maximum(np.abs(po), 1.0)).all()
    assert abs(sg - so) <= 1e-12 * max(abs(so), 1.0)
    g = ctx.grad(); go = orc.grad()
    assert (np.abs(g - go) <= 1e-12 * np.abs(go).max()).all()


def test_nlpf_general_terms(gpu, oracle_mod):
    """exponents != 1, constants, mixed sine flags, arity above the register fast path."""
    from rdis_b200 import Context
    rng = np.random.default_rng(5)
    V, F = 40, 200
    ar = rng.integers(1, 12, size=F)
    rowptr = np.concatenate([[0], np.cumsum(ar)])
    vid = np.concatenate([rng.choice(V, size=a, replace=False) for a in ar]).astype(np.int32)
    E = len(vid)
    expo = rng.choice([1.0, 2.0, 3.0, 0.5, 4.0], size=E)
    konst = rng.choice([0.0, 0.0, 0.7, -1.3], size=E)
    # keep (x-k) positive where the exponent is fractional
    sine = rng.integers(0, 2, size=E).astype(np.uint8)
    coeff = rng.normal(0, 2, size=F)
    spec = dict(kind="nlpf", V=V, F=F, lb=np.full(V, 2.0), ub=np.full(V, 5.0), rowptr=rowptr, vid=vid, expo=expo,
                konst=konst, sine=sine, coeff=coeff)
    x0 = rng.uniform(2.0, 5.0, size=V)
    ctx = Context.from_spec(spec); orc = oracle_mod.OracleFunction.from_spec(spec)
    ctx.set_x(x0); orc.set_x(x0)
    sg, pg = ctx.eval(per_factor=True); so, po = orc.eval(per_factor=True)
    assert (np.abs(pg - po) <= 1e-12 * np.maximum(np.abs(po), 1e-300)).all()
    g = ctx.grad(); go = orc.grad()
    assert (np.abs(g - go) <= 1e-11 * np.abs(go).max()).all()


def test_streaming_tile_sweep_equals_list_sweep(gpu, oracle_mod):
    """The all-factor sweeps run the streaming tile kernel (nlpf_tile_sweep.cuh), sweeps over an explicit
    factor list run the thread-per-factor kernel: per-factor values and gradients must agree to the BIT
    (same expressions), on a graph that exercises tile boundaries (arity classes of the sinusoid tree),
    an over-wide factor (more edges than a tile holds), constants, and the device-pointer entry points."""
    import torch
    from rdis_b200 import Context, problems as P
    spec = P.sinusoid(11, 2, 4)                      # V=4095, F=16364: several tiles per arity class
    # append one factor over 2100 variables (wider than kTileEdges = 2048) and a run of arity-7 factors
    rng = np.random.default_rng(11)
    wide = rng.choice(spec["V"], size=2100, replace=False).astype(np.int32)
    extra = [wide] + [rng.choice(spec["V"], size=7, replace=False).astype(np.int32) for _ in range(300)]
    lens = np.array([len(e) for e in extra])
    spec = dict(spec)
    spec["vid"] = np.concatenate([spec["vid"]] + extra).astype(np.int32)
    spec["rowptr"] = np.concatenate([spec["rowptr"], spec["rowptr"][-1] + np.cumsum(lens)])
    ne = int(lens.sum())
    spec["expo"] = np.concatenate([spec["expo"], rng.choice([1.0, 2.0, 3.0], size=ne)])
    spec["konst"] = np.concatenate([spec["konst"], rng.choice([0.0, 0.4], size=ne)])
    sn = rng.integers(0, 2, size=ne).astype(np.uint8); sn[:2100] = 1
    spec["sine"] = np.concatenate([spec["sine"], sn])
    spec["coeff"] = np.concatenate([spec["coeff"], rng.normal(0, 1, size=len(extra))])
    spec["F"] = len(spec["coeff"])
    x0 = P.random_start(spec, 5)
    ctx = Context.from_spec(spec); ctx.set_x(x0)
    allf = np.arange(spec["F"])
    for use_const in (False, True):
        if use_const:
            ctx.set_factor_const(np.array([5, 9000, spec["F"] - 1]), np.array([1.5, -2.0, 0.25]), np.ones(3, np.uint8))
        s_tile, pf_tile = ctx.eval(per_factor=True)
        s_list, pf_list = ctx.eval(allf, per_factor=True)
        assert np.array_equal(pf_tile, pf_list)
        assert abs(s_tile - s_list) <= 1e-12 * abs(s_list)
        assert np.array_equal(ctx.grad(), ctx.grad(fid=allf))
    # against the oracle (without the constant overlay)
    ctx2 = Context.from_spec(spec); ctx2.set_x(x0)
    orc = oracle_mod.OracleFunction.from_spec(spec); orc.set_x(x0)
    so, po = orc.eval(per_factor=True)
    sg, pg = ctx2.eval(per_factor=True)
    assert (np.abs(pg - po) <= 1e-12 * np.maximum(np.abs(po), 1e-300)).all()
    g = ctx2.grad(); go = orc.grad()
    assert (np.abs(g - go) <= 1e-11 * np.abs(go).max()).all()
    # device-pointer entry points: same numbers, nothing copied
    dev = torch.device("cuda", 0)
    pf_d = torch.empty(spec["F"], dtype=torch.float64, device=dev)
    tot_d = torch.zeros(1, dtype=torch.float64, device=dev)
    g_d = torch.empty(spec["V"], dtype=torch.float64, device=dev)
    ctx2.eval_device(tot_d.data_ptr(), pf_d.data_ptr())
    ctx2.grad_device(g_d.data_ptr())
    ctx2.synchronize()
    assert np.array_equal(pf_d.cpu().numpy(), pg) and float(tot_d.item()) == sg
    assert np.array_equal(g_d.cpu().numpy(), g)
    # run-to-run reproducible total
    assert ctx2.eval() == sg


def _check_solves(r, o, tol=1e-6):
    rel = _relerr(r["f_end"], o["f_end"], 1e-12)
    assert rel.max() <= tol, (rel.max(), int(rel.argmax()))
    rel0 = _relerr(r["f_init"], o["f_init"], 1e-12)
    assert rel0.max() <= 1e-12
    return rel.max()


def _oracle_batch(oracle_mod, spec, ps, x0, maxiters, variant="restated"):
    orc = oracle_mod.OracleFunction.from_spec(spec, variant)
    orc.set_x(x0)
    return orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[ps.vids], maxiters, 3e-8)


def _check_committed_state(ctx, spec, ps, x0, r):
    """bit-exact bookkeeping: returned x == committed device state; nothing else moved; the
    reported objective is the objective at the returned point; never worse than the start."""
    xg = ctx.get_x()
    assert np.array_equal(xg[ps.vids], r["x"])
    mask = np.ones(spec["V"], bool); mask[ps.vids] = False
    assert np.array_equal(xg[mask], x0[mask])
    tot = ctx.eval(ps.fids)
    assert abs(tot - r["f_end"].sum()) <= 1e-9 * abs(tot)
    assert (r["f_end"] <= r["f_init"]).all()
    assert (r["x"] >= spec["lb"][ps.vids]).all() and (r["x"] <= spec["ub"][ps.vids]).all()


def test_solve_ba_point_blocks(gpu, oracle_mod):
    """Sibling batch of 3-variable point components (sub-warp tiles)."""
    from rdis_b200 import Context, problems as P
    spec = _ba_small(P)
    x0 = spec["x0"]
    ps = P.ba_point_problems(spec)
    ctx = Context.from_spec(spec)
    ctx.set_x(x0)
    r = ctx.solve_cgd(ps, x0[ps.vids], 25, 3e-8)
    o = _oracle_batch(oracle_mod, spec, ps, x0, 25)
    worst = _check_solves(r, o, 1e-6)
    _check_committed_state(ctx, spec, ps, x0, r)
    print("point blocks: worst rel f_end diff %.3e, iters equal on %d/%d" % (worst, (r["iters"] == o["iters"]).sum(), ps.n))


def test_solve_ba_camera_blocks(gpu, oracle_mod):
    """9-variable camera components (one thread-block cluster each).  With SSmaxit = 25 these solves
    stop unconverged and are ill-conditioned: the REFERENCE's own result moves by 1e-7..3e-2 relative
    when its rounding is perturbed (same source compiled with FMA contraction: the 'twin').  No
    implementation that is not bit-identical in every libm call can promise 1e-6 there, so the contract is:
      (1) f at the start point to 1e-12, and one CG iteration from the same start to 1e-6;
      (2) at 25 iterations the GPU never returns a worse point than the start and lands within
          the band the reference itself spans under perturbation (median over problems of
          |gpu - ref| no more than 10x the median |twin - ref|, and every problem within 5e-2).
    Raising SSmaxit does not help: the unpreconditioned CG stalls on the ftol test (focal ~4e2 next to
    k2 ~1e-13 in one block) at points that differ as much as the 25-iteration ones do."""
    from rdis_b200 import Context, problems as P
    spec = _ba_small(P)
    x0 = spec["x0"]
    ps = P.ba_camera_problems(spec)
    ctx = Context.from_spec(spec)
    ctx.set_x(x0)
    r1 = ctx.solve_cgd(ps, x0[ps.vids], 1, 3e-8)
    o1 = _oracle_batch(oracle_mod, spec, ps, x0, 1)
    _check_solves(r1, o1, 1e-6)
    ctx.set_x(x0)
    r = ctx.solve_cgd(ps, x0[ps.vids], 25, 3e-8)
    _check_committed_state(ctx, spec, ps, x0, r)
    o = _oracle_batch(oracle_mod, spec, ps, x0, 25)
    try:
        of = _oracle_batch(oracle_mod, spec, ps, x0, 25, "fma")
    except Exception as e:  # host CPU without FMA
        pytest.skip("perturbation twin unavailable: %s" % e)
    spread = _relerr(of["f_end"], o["f_end"], 1e-12)
    rel = _relerr(r["f_end"], o["f_end"], 1e-12)
    print("camera blocks @25: reference-vs-twin spread", spread, "gpu-vs-reference", rel)
    assert np.median(rel) <= 10 * max(np.median(spread), 1e-7)
    assert rel.max() <= 5e-2


def test_pointwise_parity_along_reference_trajectory(gpu, oracle_mod):
    """Replay every point the ORACLE's solve evaluated (its whole CG / line-search trajectory) on the
    GPU through the C-ABI: objective to 1e-12 relative (a sum of up to ~10^2 factors in a different
    association order, with FMA contraction and CUDA's <=2-ulp sin/cos/sqrt against glibc's) and
    gradient to 1e-11 of its scale (sum of |partials| per entry) at each of them.  With the state machine proven bit-identical to the reference driver on equal inputs
    (tests/test_oracle.py::test_machine_harness_*), rounding-level evaluation noise is the only thing that can
    separate a GPU solve from a reference solve."""
    from rdis_b200 import Context, problems as P
    spec = _ba_small(P)
    x0 = spec["x0"]
    for ps in (P.ba_camera_problems(spec).subset([1]), P.ba_point_problems(spec).subset([5])):
        orc = oracle_mod.OracleFunction.from_spec(spec)
        orc.set_x(x0)
        orc.trace(True)
        orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[ps.vids], 25, 3e-8)
        recs = orc.trace_records(len(ps.vids))
        orc.trace(False)
        assert len(recs) > 20
        ctx = Context.from_spec(spec)
        ctx.set_x(x0)
        worst_f = worst_g = 0.0
        for is_df, x, out in recs[:: max(1, len(recs) // 150)]:
            ctx.set_x(x, ps.vids)
            if is_df:
                g = ctx.grad(ps.fids, ps.vids)
                # scale of a gradient entry = the sum of the magnitudes of the partials folded into it
                # (near a minimum they cancel; the entry itself can be arbitrarily small)
                rows = np.abs(ctx.factor_grad(ps.fids, 12))
                cols = rows[:, :9] if len(ps.vids) == 9 else rows[:, 9:]
                worst_g = max(worst_g, (np.abs(g - out) / np.maximum(cols.sum(axis=0), 1e-300)).max())
            else:
                f = ctx.eval(ps.fids)
                worst_f = max(worst_f, abs(f - out[0]) / abs(out[0]))
        print("trajectory replay: worst rel f %.2e, worst grad %.2e" % (worst_f, worst_g))
        assert worst_f <= 1e-12 and worst_g <= 1e-11


def test_solve_full_problem_grid(gpu, oracle_mod):
    """One large component -> cooperative-grid path (nf above the CTA threshold)."""
    from rdis_b200 import Context, problems as P
    spec = P.ba_synthetic(ncams=5, npts=40, nobs=160, seed=3)
    # enlarge so that nf > 4096 without making the oracle slow: replicate observations
    reps = 30
    spec = dict(spec)
    spec["cam"] = np.tile(spec["cam"], reps); spec["pt"] = np.tile(spec["pt"], reps)
    spec["obs"] = np.tile(spec["obs"], (reps, 1)) + np.random.default_rng(1).normal(0, 0.3, (160 * reps, 2))
    spec["F"] = 160 * reps
    x0 = spec["x0"]
    ps = P.ba_point_problems(spec)
    full = P.full_problem(spec)
    # only the points move (keeps the oracle's quadratic gradient merge cheap): one 120-var problem
    from rdis_b200.capi import ProblemSet
    big = ProblemSet([0, len(ps.vids)], ps.vids, [0, spec["F"]], np.arange(spec["F"]))
    ctx = Context.from_spec(spec); orc = oracle_mod.OracleFunction.from_spec(spec)
    ctx.set_x(x0); orc.set_x(x0)
    x0c = x0[big.vids]
    r = ctx.solve_cgd(big, x0c, 10, 3e-8)
    o = orc.solve_cgd_batch(big.var_off, big.vids, big.fac_off, big.fids, x0c, 10, 3e-8)
    _check_solves(r, o)
    assert full.n == 1


def test_solve_nlpf_subtrees(gpu, oracle_mod):
    from rdis_b200 import Context, problems as P
    spec = P.sinusoid(7, 2, 4)
    x0 = P.random_start(spec, 9)
    ps = P.sinusoid_subtree_problems(spec, 3)
    ctx = Context.from_spec(spec); orc = oracle_mod.OracleFunction.from_spec(spec)
    ctx.set_x(x0); orc.set_x(x0)
    x0c = x0[ps.vids]
    r = ctx.solve_cgd(ps, x0c, 25, 3e-8)
    o = orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0c, 25, 3e-8)
    _check_solves(r, o)


def test_edge_cases(gpu, oracle_mod):
    from rdis_b200 import Context, problems as P, RdisGpuError
    from rdis_b200.capi import ProblemSet
    spec = _ba_small(P)
    x0 = spec["x0"]
    ctx = Context.from_spec(spec)
    ctx.set_x(x0)
    # empty factor list: returns 0, delta 0, x untouched (CGD.cpp:26-29)
    ps = ProblemSet.from_lists([(np.array([63, 64, 65]), np.array([], np.int64))])
    r = ctx.solve_cgd(ps, x0[ps.vids])
    assert r["f_end"][0] == 0.0 and r["f_init"][0] == 0.0 and r["status"][0] == 5
    assert np.array_equal(r["x"], x0[ps.vids])
    # overlapping problems are refused, not silently raced
    pts = P.ba_point_problems(spec)
    bad = ProblemSet.from_lists([(pts.vids[0:3], pts.fids[pts.fac_off[0]:pts.fac_off[1]]),
                                 (pts.vids[0:3], pts.fids[pts.fac_off[1]:pts.fac_off[2]])])
    with pytest.raises(RdisGpuError):
        ctx.solve_cgd(bad, x0[bad.vids])
    # ... and so is a batch in which a factor of one problem reads a variable another problem of the batch owns (a
    # point block and the camera block that observes it are not siblings: the reference would hold one of them fixed)
    cams = P.ba_camera_problems(spec)
    c0 = int(spec["cam"][pts.fids[pts.fac_off[0]]])
    mixed = ProblemSet.from_lists([(pts.vids[0:3], pts.fids[pts.fac_off[0]:pts.fac_off[1]]),
                                   (cams.vids[9 * c0:9 * c0 + 9], cams.fids[cams.fac_off[c0]:cams.fac_off[c0 + 1]][1:])])
    with pytest.raises(RdisGpuError, match="owned by another problem"):
        ctx.solve_cgd(mixed, x0[mixed.vids])
    tree = P.sinusoid(5, 2, 4)
    tctx = Context.from_spec(tree)
    tctx.set_x(P.random_start(tree, 1))
    halves = P.sinusoid_subtree_problems(tree, 1)          # the root's two subtrees: proper siblings once the root is assigned
    tctx.solve_cgd(halves, None)
    V = tree["V"]
    allf = np.arange(tree["F"], dtype=np.int64)
    lo = np.arange(0, V // 2, dtype=np.int32); hi = np.arange(V // 2, V, dtype=np.int32)
    first = np.array([f for f in allf if tree["vid"][tree["rowptr"][f]:tree["rowptr"][f + 1]].min() < V // 2], np.int64)
    rest = np.setdiff1d(allf, first)
    torn = ProblemSet.from_lists([(lo, first), (hi, rest)])  # an arbitrary cut of the variable ids: factors straddle it
    with pytest.raises(RdisGpuError, match="owned by another problem"):
        tctx.solve_cgd(torn, None)
    # a start point outside the domain is clamped on entry (quickAssignVals, sanitize=true)
    one = pts.subset([0])
    far = x0[one.vids] + 1e9
    r = ctx.solve_cgd(one, far)
    assert (r["x"] <= spec["ub"][one.vids]).all()
    # x0 = None uses the device state
    ctx.set_x(x0)
    r1 = ctx.solve_cgd(pts.subset([1, 2]), None)
    ctx.set_x(x0)
    r2 = ctx.solve_cgd(pts.subset([1, 2]), x0[pts.subset([1, 2]).vids])
    assert np.array_equal(r1["f_end"], r2["f_end"]) and np.array_equal(r1["x"], r2["x"])


def test_assigned_constant_factors(gpu, oracle_mod):
    """Simplified factors evaluate to their constant but keep full gradients."""
    from rdis_b200 import Context, problems as P
    spec = _ba_small(P)
    x0 = spec["x0"]
    ctx = Context.from_spec(spec); orc = oracle_mod.OracleFunction.from_spec(spec)
    ctx.set_x(x0); orc.set_x(x0)
    fid = np.array([0, 5, 17]); val = np.array([1.5, -2.0, 0.25]); on = np.array([1, 1, 1], np.uint8)
    ctx.set_factor_const(fid, val, on); orc.set_factor_const(fid, val, on)
    assert abs(ctx.eval() - orc.eval()) <= 1e-12 * abs(orc.eval())
    g = ctx.grad(); go = orc.grad()
    assert (np.abs(g - go) <= 1e-11 * np.abs(go).max()).all()


def test_run_to_run_determinism(gpu):
    from rdis_b200 import Context, problems as P
    spec = _ba_small(P)
    x0 = spec["x0"]
    ctx = Context.from_spec(spec)
    ps = P.ba_point_problems(spec)
    outs = []
    for _ in range(3):
        ctx.set_x(x0)
        outs.append(ctx.solve_cgd(ps, x0[ps.vids]))
    for o in outs[1:]:
        assert np.array_equal(o["f_end"], outs[0]["f_end"]) and np.array_equal(o["x"], outs[0]["x"])
        assert np.array_equal(o["iters"], outs[0]["iters"])


def test_ba_block_kernels_match_generic_path(gpu, oracle_mod):
    """The register-resident point-block kernel and the cluster camera-block kernel against the
    generic tile / CTA kernels (`generic_only` routes a context's batches through those): same
    statuses and iteration counts where the solve is well conditioned, objectives to rounding.  Not to the bit: the
    block kernels reproduce the reference's value cache (1e-12 change filter) and, for point blocks, its summation
    order — they are BIT-IDENTICAL to the oracle's devtrig twin, asserted here too — while the generic kernels
    recompute every factor at every point and fold by trees, so the odd line search bifurcates (a handful per
    thousand blocks, like the reference itself under such a perturbation: profiles/r02_parity_attribution.json)."""
    from rdis_b200 import Context, problems as P
    spec = P.ba_synthetic(ncams=9, npts=700, nobs=3300, seed=21)
    x0 = spec["x0"]
    fast = Context.from_spec(spec)
    slow = Context.from_spec(spec)
    slow.set_option("generic_only", 1)
    fid = np.array([3, 50, 400]); val = np.array([0.5, 2.0, -1.0]); on = np.ones(3, np.uint8)
    for use_const in (False, True):
        if use_const:
            fast.set_factor_const(fid, val, on); slow.set_factor_const(fid, val, on)
        pts = P.ba_point_problems(spec)
        fast.set_x(x0); slow.set_x(x0)
        a = fast.solve_cgd(pts, x0[pts.vids], 25, 3e-8)
        b = slow.solve_cgd(pts, x0[pts.vids], 25, 3e-8)
        assert _relerr(a["f_init"], b["f_init"], 1e-12).max() <= 1e-12
        rel = _relerr(a["f_end"], b["f_end"], 1e-12)
        same = (a["iters"] == b["iters"]) & (a["status"] == b["status"])
        print("point blocks fast-vs-generic: worst rel f_end %.2e, identical iters/status on %d/%d" % (rel.max(), same.sum(), pts.n))
        assert np.median(rel) <= 1e-10 and np.quantile(rel, 0.98) <= 1e-6 and same.mean() >= 0.97
        orc = oracle_mod.OracleFunction.from_spec(spec, "devtrig")
        if use_const:
            orc.set_factor_const(fid, val, on)
        orc.set_x(x0)
        o = orc.solve_cgd_batch(pts.var_off, pts.vids, pts.fac_off, pts.fids, x0[pts.vids], 25, 3e-8)
        for key in ("f_init", "f_end", "x"):   # the production point kernel == the reference arithmetic, to the bit
            assert np.array_equal(a[key].view(np.uint64), o[key].view(np.uint64)), key
        assert np.array_equal(a["iters"], o["iters"])
        cams = P.ba_camera_problems(spec)
        fast.set_x(x0); slow.set_x(x0)
        a = fast.solve_cgd(cams, x0[cams.vids], 1, 3e-8)
        b = slow.solve_cgd(cams, x0[cams.vids], 1, 3e-8)
        assert _relerr(a["f_init"], b["f_init"]).max() <= 1e-13
        assert _relerr(a["f_end"], b["f_end"]).max() <= 1e-8
        assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["iters"], b["iters"])
        # x0 = None (device state) through the fast path
        sub = pts.subset(range(0, 700, 3))
        fast.set_x(x0); slow.set_x(x0)
        a = fast.solve_cgd(sub, None, 25, 3e-8)
        b = slow.solve_cgd(sub, None, 25, 3e-8)
        rel = _relerr(a["f_end"], b["f_end"], 1e-12)   # same band as the full batch above
        assert np.median(rel) <= 1e-10 and np.quantile(rel, 0.98) <= 1e-6
        c = fast.get_x(); d = slow.get_x()
        moved = np.zeros(spec["V"], bool); moved[sub.vids] = True
        assert np.array_equal(c[~moved], d[~moved])          # bookkeeping: nothing else moved, bit-exact
        assert np.array_equal(c[sub.vids], a["x"])


def test_golden_solves_gpu():
    """Committed golden solves on the real ladybug-49-7776 graph (written by the reference-header
    oracle build, tests/golden/make_golden.py): point blocks to 1e-6 relative on the final objective;
    camera blocks (ill-conditioned, see test_solve_ba_camera_blocks) on f_init to 1e-12 and never
    worse than the start."""
    import os
    from rdis_b200 import Context, problems as P
    from rdis_b200.capi import ProblemSet
    g = np.load(os.path.join(P.GOLDEN_DIR, "golden_solves.npz"))
    spec = P.load_golden_ba()
    ctx = Context.from_spec(spec)
    x0 = g["x0"]
    ps = ProblemSet(g["pts_var_off"], g["pts_vids"], g["pts_fac_off"], g["pts_fids"])
    ctx.set_x(x0)
    r = ctx.solve_cgd(ps, x0[ps.vids], int(g["maxiters"]), float(g["ftol"]))
    assert _relerr(r["f_init"], g["pts_f_init"]).max() <= 1e-12
    rel = _relerr(r["f_end"], g["pts_f_end"], 1e-12)
    print("golden point blocks: worst rel f_end %.3e; iters equal %d/%d" % (rel.max(), (r["iters"] == g["pts_iters"]).sum(), ps.n))
    assert rel.max() <= 1e-6
    assert abs(r["f_end"].sum() - g["pts_f_end"].sum()) <= 1e-6 * g["pts_f_end"].sum()
    ps = ProblemSet(g["cams_var_off"], g["cams_vids"], g["cams_fac_off"], g["cams_fids"])
    ctx.set_x(x0)
    r = ctx.solve_cgd(ps, x0[ps.vids], int(g["maxiters"]), float(g["ftol"]))
    assert _relerr(r["f_init"], g["cams_f_init"]).max() <= 1e-12
    assert (r["f_end"] <= r["f_init"]).all()
    print("golden camera blocks: gpu f_end", r["f_end"], "reference", g["cams_f_end"])
    # the file's own initial state: full objective
    ctx.set_x(spec["x0"])
    assert abs(ctx.eval() - float(g["f_file_x0"])) <= 1e-12 * float(g["f_file_x0"])


def test_struct_and_csr_entry_points_agree(gpu):
    """rdisgpu_solve_cgd (array of rdisgpu_problem structs, per-problem x0 pointers, some NULL = device
    state) and rdisgpu_solve_cgd_csr (packed lists) are the same solve: identical to the bit."""
    from rdis_b200 import Context, problems as P
    spec = _ba_small(P)
    x0 = spec["x0"]
    ctx = Context.from_spec(spec)
    for ps in (P.ba_point_problems(spec), P.ba_camera_problems(spec), P.full_problem(spec)):
        ctx.set_x(x0)
        a = ctx.solve_cgd(ps, x0[ps.vids], 5, 3e-8)
        ctx.set_x(x0)
        b = ctx.solve_cgd_structs(ps, x0[ps.vids], 5, 3e-8)
        for key in ("f_init", "f_end", "x", "iters", "status", "n_feval", "n_geval"):
            assert np.array_equal(a[key], b[key]), key


@pytest.mark.gpu
def test_solve_lm_blocks(gpu, oracle_mod):
    """Levenberg-Marquardt subspace solves (rdisgpu_solve_lm_csr) against the oracle's restatement of
    LMSubspaceOptimizer + levmar's dlevmar_der.  PARITY UNPINNED upstream (levmar is not vendored and no
    reference test runs LM): this pins the device path to OUR restatement only.  Point blocks (3 x 3 normal
    equations) and camera blocks (9 x 9, 361..906 residuals): same iteration counts and stop codes, final
    objective to 1e-6 relative, start objective to 1e-12; and the bookkeeping contract of the boundary."""
    from rdis_b200 import Context, problems as P
    spec = P.ba_synthetic(ncams=6, npts=150, nobs=640, seed=12)
    x0 = spec["x0"]
    ctx = Context.from_spec(spec)
    for name, ps, iters in (("points", P.ba_point_problems(spec), 25), ("cameras", P.ba_camera_problems(spec), 8)):
        ctx.set_x(x0)
        r = ctx.solve_lm(ps, x0[ps.vids], iters, 3e-8)
        orc = oracle_mod.OracleFunction.from_spec(spec); orc.set_x(x0)
        o = orc.solve_lm_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[ps.vids], iters, 3e-8)
        rel0 = _relerr(r["f_init"], o["f_init"], 1e-12)
        rel = _relerr(r["f_end"], o["f_end"], 1e-12)
        same = (r["iters"] == o["iters"]) & (r["stop"] == o["stop"])
        print("LM %s: worst rel f_init %.2e f_end %.2e; identical iters/stop on %d/%d; stop histogram %s" %
              (name, rel0.max(), rel.max(), same.sum(), ps.n, np.bincount(r["stop"], minlength=8).tolist()))
        assert rel0.max() <= 1e-12
        assert rel.max() <= 1e-6
        assert same.mean() >= 0.97
        # boundary bookkeeping: returned x is the committed state, nothing else moved, objective consistent
        xg = ctx.get_x()
        assert np.array_equal(xg[ps.vids], r["x"])
        mask = np.ones(spec["V"], bool); mask[ps.vids] = False
        assert np.array_equal(xg[mask], x0[mask])
        tot = ctx.eval(ps.fids)
        assert abs(tot - r["f_end"].sum()) <= 1e-9 * abs(tot)
        assert (r["f_end"] <= r["f_init"] * (1 + 1e-12)).all()
    # a component of more than 32 variables takes the dense solver (lm_dense.cuh): the whole graph as ONE problem
    # (504 variables, 640 residuals; the oracle's dense levmar restatement needs ~1 s for it)
    big = P.full_problem(spec)
    ctx.set_x(x0)
    r = ctx.solve_lm(big, x0[big.vids], 6, 3e-8)
    orc = oracle_mod.OracleFunction.from_spec(spec); orc.set_x(x0)
    o = orc.solve_lm_batch(big.var_off, big.vids, big.fac_off, big.fids, x0[big.vids], 6, 3e-8)
    print("LM whole graph (m=%d): f %.6e -> gpu %.9e oracle %.9e, iters %d/%d stop %d/%d" % (
        len(big.vids), r["f_init"][0], r["f_end"][0], o["f_end"][0], r["iters"][0], o["iters"][0], r["stop"][0], o["stop"][0]))
    assert abs(r["f_init"][0] - o["f_init"][0]) <= 1e-12 * abs(o["f_init"][0])
    assert abs(r["f_end"][0] - o["f_end"][0]) <= 1e-6 * abs(o["f_end"][0])
    assert r["iters"][0] == o["iters"][0] and r["stop"][0] == o["stop"][0]
    assert np.array_equal(ctx.get_x()[big.vids], r["x"]) and r["f_end"][0] < r["f_init"][0]
    # an empty component: contract of the boundary
    emp = gpu.ProblemSet.from_lists([(np.array([9 * 6, 9 * 6 + 1, 9 * 6 + 2], np.int32), np.zeros(0, np.int64))])
    ctx.set_x(x0)
    r = ctx.solve_lm(emp, x0[emp.vids], 5, 3e-8)
    assert r["f_end"][0] == 0 and np.array_equal(r["x"], x0[emp.vids])


def test_solve_lm_dense_components(gpu, oracle_mod):
    """Levenberg-Marquardt on components of more than 32 variables (lm_dense.cuh: block-sparse Jacobian rows, dense
    normal equations assembled per variable, blocked Cholesky with the trailing update on the FP64 tensor cores)
    against the oracle's dense levmar restatement: sinusoid subtrees (127 / 255 variables; several per call, sizes that
    are not multiples of the 64-wide blocks) and a two-camera bundle-adjustment block.  PARITY UNPINNED upstream."""
    from rdis_b200 import Context, problems as P
    from rdis_b200.problems import ProblemSet
    tree = P.sinusoid(8, 2, 4)
    xt = P.random_start(tree, 9)
    ctx = Context.from_spec(tree)
    orc = oracle_mod.OracleFunction.from_spec(tree)
    for lv, iters in ((2, 12), (1, 8)):
        ps = P.sinusoid_subtree_problems(tree, lv)
        ctx.set_x(xt); orc.set_x(xt)
        r = ctx.solve_lm(ps, xt[ps.vids], iters, 3e-8)
        o = orc.solve_lm_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, xt[ps.vids], iters, 3e-8)
        rel = _relerr(r["f_end"], o["f_end"], 1e-12)
        print("LM dense, %d subtrees of %d variables: worst rel f_end %.2e, iters %s / %s, stop %s / %s" % (
            ps.n, ps.var_off[1], rel.max(), r["iters"].tolist(), o["iters"].tolist(), r["stop"].tolist(), o["stop"].tolist()))
        assert _relerr(r["f_init"], o["f_init"], 1e-12).max() <= 1e-12
        assert rel.max() <= 1e-6
        assert np.array_equal(r["iters"], o["iters"]) and np.array_equal(r["stop"], o["stop"])
        assert np.array_equal(ctx.get_x()[ps.vids], r["x"])
    spec = P.ba_synthetic(ncams=4, npts=90, nobs=330, seed=2)
    x0 = spec["x0"]
    cam_sel = np.array([0, 2]); pt_sel = np.unique(spec["pt"][np.isin(spec["cam"], cam_sel)])[:40]
    vids = np.sort(np.concatenate([(9 * cam_sel[:, None] + np.arange(9)).ravel(), (9 * 4 + 3 * pt_sel[:, None] + np.arange(3)).ravel()])).astype(np.int32)
    fids = np.nonzero(np.isin(spec["cam"], cam_sel) | np.isin(spec["pt"], pt_sel))[0].astype(np.int64)
    ps = ProblemSet([0, len(vids)], vids, [0, len(fids)], fids)
    bctx = Context.from_spec(spec); borc = oracle_mod.OracleFunction.from_spec(spec)
    bctx.set_x(x0); borc.set_x(x0)
    r = bctx.solve_lm(ps, x0[vids], 10, 3e-8)
    o = borc.solve_lm_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[vids], 10, 3e-8)
    print("LM dense, BA block m=%d n=%d: f %.6e -> gpu %.9e oracle %.9e, iters %d/%d stop %d/%d" % (
        len(vids), len(fids), r["f_init"][0], r["f_end"][0], o["f_end"][0], r["iters"][0], o["iters"][0], r["stop"][0], o["stop"][0]))
    assert abs(r["f_end"][0] - o["f_end"][0]) <= 1e-6 * abs(o["f_end"][0]) and r["iters"][0] == o["iters"][0]


@pytest.mark.gpu
def test_ba_rows_sweep_matches_list_path(gpu, oracle_mod):
    """rdisgpu_factor_rows_device (the pipelined all-factor residual + Jacobian-rows sweep) against the per-factor list
    path (rdisgpu_factor_grad / rdisgpu_eval, bit for bit: same expressions) and the oracle (1e-11), on a graph with
    assigned-constant factors; the sweep's total against the sum of its values."""
    import torch
    from rdis_b200 import Context, problems as P
    spec = P.ba_synthetic(ncams=11, npts=900, nobs=4001, seed=5)
    ctx = Context.from_spec(spec)
    ctx.set_x(spec["x0"])
    fid = np.array([0, 17, 4000], np.int64)
    ctx.set_factor_const(fid, np.array([1.5, -2.0, 0.25]), np.ones(3, np.uint8))
    F = spec["F"]
    pf = torch.zeros(F, dtype=torch.float64, device="cuda")
    rows = torch.zeros(F * 12, dtype=torch.float64, device="cuda")
    tot = torch.zeros(1, dtype=torch.float64, device="cuda")
    ctx.factor_rows_device(pf.data_ptr(), rows.data_ptr(), tot.data_ptr())
    torch.cuda.synchronize()
    _, want_pf = ctx.eval(per_factor=True)
    want_rows = ctx.factor_grad(np.arange(F), 12)
    assert np.array_equal(pf.cpu().numpy().view(np.uint64), want_pf.view(np.uint64))
    assert np.array_equal(rows.cpu().numpy().reshape(F, 12).view(np.uint64), want_rows.view(np.uint64))
    assert abs(float(tot.item()) - want_pf.sum()) <= 1e-12 * abs(want_pf.sum())
    orc = oracle_mod.OracleFunction.from_spec(spec)
    orc.set_x(spec["x0"])
    ref = np.stack([orc.factor_grad(j, 12) for j in range(0, F, 7)])
    got = rows.cpu().numpy().reshape(F, 12)[::7]
    assert (np.abs(got - ref) <= 1e-11 * np.abs(ref).max(axis=1, keepdims=True) + 1e-300).all()


def test_corrected_quotients_are_exact(built_lib):
    """QuotBy<kCorrected> (one reciprocal + a correction step per quotient: what the production BA gradient uses for the
    reference's 24 divisions per observation) returns the division's own bits on 2^28 operand pairs."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "rdis_b200", "host"), "all"], stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(root, "tests", "native", "quot_check")], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr


def test_inline_trig_is_bit_identical(built_lib):
    """rdis_sin / rdis_cos / rdis_sincos (the inline kernels the NLPF terms use) == the CUDA library's
    sin / cos to the bit on 8 M values incl. the slow-path threshold and the special values."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "rdis_b200", "host"), "all"], stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(root, "tests", "native", "trig_check")], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_more_edge_cases(gpu, oracle_mod):
    """Clamping active inside the line search, a stationary start, a ragged batch that mixes block shapes, and
    the error paths of the C-ABI (bad ids, wrong call order) — against the oracle where there is a number."""
    from rdis_b200 import Context, problems as P, RdisGpuError
    from rdis_b200.capi import ProblemSet
    # (1) optimum outside the domain: 0.5*(x-1)^2 + 0.5*(y+3)^2 on [2,5] x [-2,2]  ->  clamps to (2, -2)
    V = 2
    spec = dict(kind="nlpf", V=V, F=2, lb=np.array([2.0, -2.0]), ub=np.array([5.0, 2.0]), rowptr=np.array([0, 1, 2]),
                vid=np.array([0, 1], np.int32), expo=np.array([2.0, 2.0]), konst=np.array([1.0, -3.0]),
                sine=np.zeros(2, np.uint8), coeff=np.array([0.5, 0.5]))
    ctx = Context.from_spec(spec)
    orc = oracle_mod.OracleFunction.from_spec(spec)
    ps = ProblemSet.from_lists([(np.array([0, 1], np.int32), np.array([0, 1], np.int64))])
    for start in (np.array([3.0, 0.5]), np.array([2.0, -2.0]), np.array([4.9, 1.9])):
        ctx.set_x(start); orc.set_x(start)
        r = ctx.solve_cgd(ps, start, 25, 3e-8)
        o = orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, start, 25, 3e-8)
        assert np.allclose(r["x"], o["x"], rtol=0, atol=1e-12) and abs(r["f_end"][0] - o["f_end"][0]) <= 1e-12 * abs(o["f_end"][0])
        assert r["iters"][0] == o["iters"][0]
        assert np.allclose(r["x"], [2.0, -2.0], atol=1e-9)
    # (2) a stationary start: the gradient test returns at once, nothing moves
    spec2 = dict(spec); spec2["lb"] = np.array([-9.0, -9.0]); spec2["ub"] = np.array([9.0, 9.0])
    ctx2 = Context.from_spec(spec2); orc2 = oracle_mod.OracleFunction.from_spec(spec2)
    start = np.array([1.0, -3.0])
    ctx2.set_x(start); orc2.set_x(start)
    r = ctx2.solve_cgd(ps, start, 25, 3e-8)
    o = orc2.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, start, 25, 3e-8)
    assert r["f_end"][0] == 0.0 == o["f_end"][0] and np.array_equal(r["x"], start) and r["iters"][0] == o["iters"][0]
    # (3) ragged batch: two camera blocks + the point blocks those cameras do not observe, in ONE call
    sp = P.ba_synthetic(ncams=6, npts=80, nobs=300, seed=17)
    x0 = sp["x0"]
    cams, pts = P.ba_camera_problems(sp), P.ba_point_problems(sp)
    seen = set(sp["pt"][(sp["cam"] == 0) | (sp["cam"] == 1)].tolist())
    free_pts = [i for i in range(80) if i not in seen]
    assert len(free_pts) >= 5
    probs = [(cams.vids[cams.var_off[c]:cams.var_off[c + 1]], cams.fids[cams.fac_off[c]:cams.fac_off[c + 1]]) for c in (0, 1)]
    probs += [(pts.vids[pts.var_off[i]:pts.var_off[i + 1]], pts.fids[pts.fac_off[i]:pts.fac_off[i + 1]]) for i in free_pts]
    mixed = ProblemSet.from_lists(probs)
    c3 = Context.from_spec(sp); o3 = oracle_mod.OracleFunction.from_spec(sp)
    c3.set_x(x0); o3.set_x(x0)
    r = c3.solve_cgd(mixed, x0[mixed.vids], 1, 3e-8)      # one CG iteration: stable for the camera blocks too
    o = o3.solve_cgd_batch(mixed.var_off, mixed.vids, mixed.fac_off, mixed.fids, x0[mixed.vids], 1, 3e-8)
    assert _relerr(r["f_init"], o["f_init"], 1e-12).max() <= 1e-12
    assert _relerr(r["f_end"], o["f_end"], 1e-12).max() <= 1e-6
    # (4) error paths: loud, with a message, and the context stays usable
    with pytest.raises(RdisGpuError):
        c3.solve_cgd(ProblemSet.from_lists([(np.array([sp["V"]], np.int32), np.array([0], np.int64))]), np.zeros(1))
    with pytest.raises(RdisGpuError):
        c3.solve_cgd(ProblemSet.from_lists([(np.array([0], np.int32), np.array([sp["F"]], np.int64))]), np.zeros(1))
    with pytest.raises(RdisGpuError):
        c3.set_x(np.zeros(3), vid=np.array([0, 1, sp["V"] + 5], np.int32))
    with pytest.raises(RdisGpuError):
        c3.solve_cgd(mixed, x0[mixed.vids], 0, 3e-8)       # maxiters must be positive
    fresh = Context()
    with pytest.raises(RdisGpuError):
        fresh.eval()                                        # before finalize
    c3.set_x(x0); o3.set_x(x0)
    assert abs(c3.eval() - o3.eval()) <= 1e-12 * abs(o3.eval())


@pytest.mark.gpu
def test_ba_table_sweep_equals_list_sweep(gpu, oracle_mod):
    """The all-factor BA sweep (per-camera table + streaming kernel, ba_sweep.cuh) against the thread-per-factor
    list sweep: per-factor values to the BIT, incl. assigned-constant factors, a zero rotation vector (the theta == 0
    branch), after solves have moved the state, and through the device-pointer entry; and against the oracle."""
    import torch
    from rdis_b200 import Context, problems as P
    spec = P.ba_synthetic(ncams=9, npts=400, nobs=1900, seed=23)
    x0 = spec["x0"].copy()
    x0[0:3] = 0.0                                   # camera 0: theta == 0
    ctx = Context.from_spec(spec); ctx.set_x(x0)
    allf = np.arange(spec["F"])
    s_tab, pf_tab = ctx.eval(per_factor=True)
    s_lst, pf_lst = ctx.eval(allf, per_factor=True)
    assert np.array_equal(pf_tab, pf_lst) and abs(s_tab - s_lst) <= 1e-12 * abs(s_lst)
    orc = oracle_mod.OracleFunction.from_spec(spec); orc.set_x(x0)
    so, po = orc.eval(per_factor=True)
    assert (np.abs(pf_tab - po) <= 1e-12 * np.maximum(np.abs(po), 1e-300)).all()
    ctx.set_factor_const(np.array([7, 1000]), np.array([0.25, 3.0]), np.ones(2, np.uint8))
    assert np.array_equal(ctx.eval(per_factor=True)[1], ctx.eval(allf, per_factor=True)[1])
    pts = P.ba_point_problems(spec)
    ctx.solve_cgd(pts, x0[pts.vids], 5, 3e-8)       # commits new point values: the value mirror must follow
    pf_a = ctx.eval(per_factor=True)[1]; pf_b = ctx.eval(allf, per_factor=True)[1]
    assert np.array_equal(pf_a, pf_b)
    dev = torch.device("cuda", 0)
    pf_d = torch.empty(spec["F"], dtype=torch.float64, device=dev); tot = torch.zeros(1, dtype=torch.float64, device=dev)
    ctx.eval_device(tot.data_ptr(), pf_d.data_ptr()); ctx.synchronize()
    assert np.array_equal(pf_d.cpu().numpy(), pf_a)


def _scipy_labels(spec, assigned, fconst=()):
    """Canonical (min variable id) component labels of the reference's connectivity rule, computed independently."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    V, F = spec["V"], spec["F"]
    if spec["kind"] == "nlpf":
        lens = np.diff(spec["rowptr"]); fe = np.repeat(np.arange(F), lens); ve = spec["vid"].astype(np.int64)
    else:
        nc = spec["ncams"]
        fe = np.repeat(np.arange(F), 12)
        ve = np.concatenate([9 * spec["cam"].astype(np.int64)[:, None] + np.arange(9), 9 * nc + 3 * spec["pt"].astype(np.int64)[:, None] + np.arange(3)], axis=1).ravel()
    dead = np.zeros(F, bool); dead[list(fconst)] = True
    keep = (~assigned.astype(bool)[ve]) & (~dead[fe])
    g = coo_matrix((np.ones(keep.sum()), (ve[keep], V + fe[keep])), shape=(V + F, V + F))
    _, lab = connected_components(g, directed=False)
    vl = np.full(V, -1, np.int64)
    un = np.nonzero(assigned == 0)[0]
    first = {}
    for v in un:                       # ascending: the first variable seen in a component is its smallest id
        first.setdefault(lab[v], v)
        vl[v] = first[lab[v]]
    fl = np.full(F, -1, np.int64)
    has = np.zeros(F, bool); np.logical_or.at(has, fe[keep], True)
    fl[has] = np.array([first[lab[V + f]] for f in np.nonzero(has)[0]])
    return vl, fl


@pytest.mark.gpu
def test_device_components_membership_is_exact(gpu, oracle_mod):
    """rdisgpu_components (min-label propagation on the device) against an independent scipy labelling of the same
    connectivity rule — integer bookkeeping: exact — on BA cuts, the sinusoid tree and chain (diameter 1000), random
    assignments and assigned-constant factors; and the packed ProblemSet against the generators the bench uses."""
    from rdis_b200 import Context, problems as P
    rng = np.random.default_rng(3)
    ba = P.ba_synthetic(ncams=7, npts=300, nobs=1300, seed=41)
    tree = P.sinusoid(9, 2, 4)
    chain = P.sinusoid(999, 1, 3)
    cases = []
    a = np.zeros(ba["V"], np.uint8); a[:63] = 1; cases.append((ba, a, ()))
    a = np.zeros(ba["V"], np.uint8); a[63:] = 1; cases.append((ba, a, ()))
    a = (rng.random(ba["V"]) < 0.5).astype(np.uint8); cases.append((ba, a, (5, 77, 900)))
    a = np.zeros(tree["V"], np.uint8); a[:15] = 1; cases.append((tree, a, ()))
    a = (rng.random(tree["V"]) < 0.3).astype(np.uint8); cases.append((tree, a, (0, 100, 2000)))
    a = np.zeros(chain["V"], np.uint8); cases.append((chain, a, ()))
    a = np.zeros(chain["V"], np.uint8); a[::97] = 1; cases.append((chain, a, ()))
    ctxs = {}
    for sp, a, fc in cases:
        ctx = Context.from_spec(sp)
        ctx.set_x(np.zeros(sp["V"]))
        if fc:
            ctx.set_factor_const(np.array(fc), np.zeros(len(fc)), np.ones(len(fc), np.uint8))
        vl, fl, n, rounds = ctx.components(a)
        wv, wf = _scipy_labels(sp, a, fc)
        assert np.array_equal(vl, wv) and np.array_equal(fl, wf)
        assert n == len(np.unique(wv[wv >= 0]))
        print("components: V=%d F=%d -> %d components in %d rounds" % (sp["V"], sp["F"], n, rounds))
    # the packed problem set == what the generators produce for the same cut
    ctx = Context.from_spec(ba)
    a = np.zeros(ba["V"], np.uint8); a[:63] = 1
    got = ctx.component_problems(a); want = P.ba_point_problems(ba)
    assert np.array_equal(got.var_off, want.var_off) and np.array_equal(got.vids, want.vids)
    assert np.array_equal(got.fac_off, want.fac_off) and np.array_equal(got.fids, want.fids)
    ctx = Context.from_spec(tree)
    a = np.zeros(tree["V"], np.uint8); a[:7] = 1
    got = ctx.component_problems(a); want = P.sinusoid_subtree_problems(tree, 3)
    assert np.array_equal(got.vids, want.vids) and np.array_equal(got.fids, want.fids) and np.array_equal(got.fac_off, want.fac_off)


def _two_group_nlpf(seed, V=60, F=260, own=20):
    """Random NonlinearProductFactor graph with two sibling components (variables [0, own) and [own, 2*own)) over a
    pool of frozen variables [2*own, V): mixed exponents / constants / sine flags, repeated (k, e, sine) expressions
    on one variable (shared terms), arities up to 7 (above the register fast path)."""
    rng = np.random.default_rng(seed)
    rows, groups = [], []
    for j in range(F):
        g = int(rng.integers(0, 2))
        a = int(rng.integers(1, 8))
        n_own = int(rng.integers(1, a + 1))
        vs = list(rng.choice(np.arange(g * own, (g + 1) * own), size=min(n_own, own), replace=False))
        vs += list(rng.choice(np.arange(2 * own, V), size=a - len(vs), replace=False))
        rng.shuffle(vs)
        rows.append(np.array(vs, np.int32)); groups.append(g)
    rowptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
    vid = np.concatenate(rows).astype(np.int32)
    E = len(vid)
    expo = rng.choice([1.0, 1.0, 2.0, 3.0], size=E)
    konst = rng.choice([0.0, 0.0, 0.7, -1.3], size=E)
    sine = rng.integers(0, 2, size=E).astype(np.uint8)
    coeff = rng.normal(0, 0.5, size=F)
    spec = dict(kind="nlpf", V=V, F=F, lb=np.full(V, -3.0), ub=np.full(V, 3.0), rowptr=rowptr, vid=vid, expo=expo,
                konst=konst, sine=sine, coeff=coeff)
    groups = np.array(groups)
    probs = [(np.arange(g * own, (g + 1) * own, dtype=np.int32), np.nonzero(groups == g)[0].astype(np.int64)) for g in (0, 1)]
    return spec, probs, rng.uniform(-2.5, 2.5, size=V)


@pytest.mark.gpu
def test_nlpf_resident_kernel_matches_generic_path(gpu, oracle_mod):
    """The shared-memory resident NonlinearProductFactor component kernel (distinct terms evaluated once per line
    point) against the generic CTA kernel (`generic_only`).  At 256 threads it folds the same per-factor expressions
    over the same factor -> thread mapping and reduction tree, so the whole result is demanded EQUAL (objective,
    iterations, status, evaluation counts, committed state); the default 512-thread build sums in another order and
    is held to the oracle tolerance.  All within 1e-6 of the CPU oracle."""
    from rdis_b200 import Context, problems as P
    from rdis_b200.capi import ProblemSet
    cases = []
    tree = P.sinusoid(9, 2, 4)
    cases.append(("sinusoid h=9 subtrees", tree, P.sinusoid_subtree_problems(tree, 3), P.random_start(tree, 11), ()))
    big = P.sinusoid(12, 2, 4)   # components of BASELINE config 4's shape: 1023 variables / 4092 factors, 3 frozen ancestors
    cases.append(("config-4-shaped components", big, P.sinusoid_subtree_problems(big, 3).subset([0, 5]), P.random_start(big, 4), ()))
    spec, probs, x0 = _two_group_nlpf(17)
    cases.append(("two-group general terms", spec, ProblemSet.from_lists(probs), x0, ()))
    cases.append(("two-group + assigned constants", spec, ProblemSet.from_lists(probs), x0, (3, 40, 41, 200)))
    for name, sp, ps, x0, fconst in cases:
        fast = Context.from_spec(sp); slow = Context.from_spec(sp); wide = Context.from_spec(sp)
        slow.set_option("generic_only", 1)
        fast.set_option("resident_threads", 256)
        orc = oracle_mod.OracleFunction.from_spec(sp)
        if fconst:
            fid = np.array(fconst); val = np.linspace(-1.0, 2.0, len(fid)); on = np.ones(len(fid), np.uint8)
            for c in (fast, slow, wide, orc):
                c.set_factor_const(fid, val, on)
        fast.set_x(x0); slow.set_x(x0); orc.set_x(x0); wide.set_x(x0)
        x0c = x0[ps.vids]
        bw = wide.batch(ps); winfo = bw.info(); bw.close()
        assert winfo["resident_problems"] == ps.n
        w = wide.solve_cgd(ps, x0c, 25, 3e-8)
        bf = fast.batch(ps); info = bf.info(); bf.close()
        assert info["resident_problems"] == ps.n and info["generic_problems"] == 0, info
        a = fast.solve_cgd(ps, x0c, 25, 3e-8)
        b = slow.solve_cgd(ps, x0c, 25, 3e-8)
        o = orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0c, 25, 3e-8)
        rel = _relerr(a["f_end"], b["f_end"], 1e-12)
        exact = all(np.array_equal(a[k], b[k]) for k in ("f_init", "f_end", "iters", "status", "x"))
        print("%s: %d problems, resident smem %d B, worst rel f_end vs generic %.2e, exact=%s, evals %d" % (
            name, ps.n, info["resident_smem_bytes"], rel.max(), exact, int(a["n_feval"].sum() + a["n_geval"].sum())))
        assert exact
        assert np.array_equal(a["n_feval"], b["n_feval"]) and np.array_equal(a["n_geval"], b["n_geval"])
        assert np.array_equal(fast.get_x(), slow.get_x())
        _check_solves(a, o)
        _check_solves(w, o)
        assert np.array_equal(w["f_init"], a["f_init"]) or _relerr(w["f_init"], a["f_init"], 1e-12).max() <= 1e-13
        _check_committed_state(fast, sp, ps, x0, a)
        _check_committed_state(wide, sp, ps, x0, w)
        # device-state start (x0 = None) on a subset: nothing outside the subset moves
        sub = ps.subset(range(0, ps.n, 2))
        fast.set_x(x0); slow.set_x(x0)
        a = fast.solve_cgd(sub, None, 25, 3e-8); b = slow.solve_cgd(sub, None, 25, 3e-8)
        assert np.array_equal(a["f_end"], b["f_end"]) and np.array_equal(a["x"], b["x"])
        assert np.array_equal(fast.get_x(), slow.get_x())


@pytest.mark.gpu
def test_config2_whole_graph_as_one_problem(gpu, oracle_mod):
    """BASELINE config 2 (optSinusoid d=1000): every variable and factor of the chain (h=999, k=1) and of the default-
    shaped tree (h=6, k=3) in ONE subspace problem — no frozen variables, the resident kernel's widest case.  Equal to
    the generic CTA kernel at 256 threads; within the band of the reference's own rounding twin against the oracle
    (a 1000-variable, 25-iteration CG on a sum of sines is not a contraction: see test_solve_ba_camera_blocks)."""
    from rdis_b200 import Context, problems as P
    for h, k, ar in ((999, 1, 3), (6, 3, 3)):
        sp = P.sinusoid(h, k, ar)
        x0 = P.random_start(sp, 727)
        ps = P.full_problem(sp)
        fast = Context.from_spec(sp); slow = Context.from_spec(sp); wide = Context.from_spec(sp)
        slow.set_option("generic_only", 1); fast.set_option("resident_threads", 256)
        for c in (fast, slow, wide):
            c.set_x(x0)
        bi = wide.batch(ps); info = bi.info(); bi.close()
        assert info["resident_problems"] == 1, info
        a = fast.solve_cgd(ps, x0, 25, 3e-8); b = slow.solve_cgd(ps, x0, 25, 3e-8); w = wide.solve_cgd(ps, x0, 25, 3e-8)
        assert all(np.array_equal(a[key], b[key]) for key in ("f_init", "f_end", "iters", "status", "x", "n_feval", "n_geval"))
        o = _oracle_batch(oracle_mod, sp, ps, x0, 25)
        assert _relerr(a["f_init"], o["f_init"], 1e-12).max() <= 1e-12 and _relerr(w["f_init"], o["f_init"], 1e-12).max() <= 1e-12
        try:
            of = _oracle_batch(oracle_mod, sp, ps, x0, 25, "fma")
            spread = float(_relerr(of["f_end"], o["f_end"], 1e-12).max())
        except Exception:
            spread = 0.0
        rel = max(float(_relerr(a["f_end"], o["f_end"], 1e-12).max()), float(_relerr(w["f_end"], o["f_end"], 1e-12).max()))
        print("config 2 h=%d k=%d: V=%d F=%d, gpu-vs-oracle rel f_end %.2e, oracle-vs-FMA-twin %.2e, evals %d" % (
            h, k, sp["V"], sp["F"], rel, spread, int(a["n_feval"][0] + a["n_geval"][0])))
        assert rel <= max(1e-6, 10.0 * spread)
        assert (w["f_end"] <= w["f_init"]).all() and (a["f_end"] <= a["f_init"]).all()


def test_launch_order_history_does_not_change_results(gpu):
    """The block kernels re-sort their launch order after every solve (longest chains of the previous visit first,
    rdisgpu_set_option "adaptive_order"); a problem is solved by its own cluster / tile whatever its position, so three
    visits of the real ladybug wave with the history on and off agree to the bit, problem by problem."""
    from rdis_b200 import Context, problems as P
    spec = P.load_golden_ba()
    x0 = spec["x0"]
    pts, cams = P.ba_point_problems(spec), P.ba_camera_problems(spec)
    out = {}
    for adaptive in (0, 1):
        ctx = Context.from_spec(spec)
        ctx.set_option("adaptive_order", adaptive)
        bp, bc = ctx.batch(pts), ctx.batch(cams)
        visits = []
        for visit in range(3):
            ctx.set_x(x0)
            bp.solve(x0[pts.vids].copy(), 25, 3e-8)
            rp = bp.fetch()
            bc.solve(None, 25, 3e-8)
            rc = bc.fetch()
            visits.append((rp, rc))
        out[adaptive] = visits
        ctx.close()
    for visit in range(3):
        for side in (0, 1):
            a, b = out[0][visit][side], out[1][visit][side]
            for key in ("x", "f_init", "f_end"):
                assert np.array_equal(a[key].view(np.uint64), b[key].view(np.uint64)), (visit, side, key)
            for key in ("iters", "status", "n_feval"):
                assert np.array_equal(a[key], b[key]), (visit, side, key)
    # and every visit reproduces the first one (same start state)
    for side in (0, 1):
        assert np.array_equal(out[1][0][side]["f_end"].view(np.uint64), out[1][2][side]["f_end"].view(np.uint64))
