"""Strict mode (rdisgpu_set_option "strict") against the oracle's devtrig twin: BIT IDENTITY.

The devtrig twin is the CPU restatement with sin/cos taken from rdis_b200/csrc/trig.cuh compiled for the host
(oracle/Makefile).  Every other operation of the path is a correctly rounded IEEE operation on both sides, so the
strict kernel — the reference's operation order, value cache with its 1e-12 change filter, evaluation sequence —
must return the same BITS: final point, f_init, f_end, iteration count, on every component.  These tests are the
proof that the device implements CGDSubspaceOptimizer::optimize (src/optimizers/CGDSubspaceOptimizer.cpp:19-98)
and not something near it; the fast kernels are then compared with the strict kernel / the oracle at 1e-6
(tests/test_gpu_parity.py) with every remaining difference attributed.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def _assert_identical(r, o, what):
    for key in ("f_end", "f_init", "x"):
        bad = np.nonzero(_bits(r[key]) != _bits(o[key]))[0]
        assert bad.size == 0, "%s: %s differs at %d entries, first %d: gpu %r cpu %r" % (
            what, key, bad.size, bad[0], r[key][bad[0]], o[key][bad[0]])
    assert np.array_equal(r["iters"], o["iters"]), what + ": iteration counts differ"


@pytest.fixture(scope="module")
def ladybug(built_lib, oracle_mod):
    from rdis_b200 import problems as P
    spec = P.load_golden_ba()
    return spec, P.ba_point_problems(spec), P.ba_camera_problems(spec)


def _pair(spec, oracle_mod, x):
    from rdis_b200 import Context
    ctx = Context.from_spec(spec)
    ctx.set_option("strict", 1)
    ctx.set_x(x)
    orc = oracle_mod.OracleFunction.from_spec(spec, "devtrig")
    orc.set_x(x)
    return ctx, orc


def test_strict_full_ladybug_step_is_bit_identical(ladybug, oracle_mod):
    """BASELINE config 3 from the file's own state: ALL 7776 point components, then ALL 49 camera components on the
    state the point wave left (value caches persisting across the two calls on both sides), and the step's
    objective — identical bits."""
    spec, pts, cams = ladybug
    x0 = spec["x0"]
    ctx, orc = _pair(spec, oracle_mod, x0)
    r = ctx.solve_cgd(pts, x0[pts.vids], 25, 3e-8)
    o = orc.solve_cgd_batch(pts.var_off, pts.vids, pts.fac_off, pts.fids, x0[pts.vids], 25, 3e-8)
    _assert_identical(r, o, "point wave")
    x1 = ctx.get_x()
    assert np.array_equal(_bits(x1), _bits(orc.get_x()))
    r2 = ctx.solve_cgd(cams, x1[cams.vids], 25, 3e-8)
    o2 = orc.solve_cgd_batch(cams.var_off, cams.vids, cams.fac_off, cams.fids, x1[cams.vids], 25, 3e-8)
    _assert_identical(r2, o2, "camera wave after the point wave")
    assert float(r2["f_end"].sum()) == float(o2["f_end"].sum())
    assert np.array_equal(_bits(ctx.get_x()), _bits(orc.get_x()))
    # evaluation counts: the device requests exactly the evaluations the reference performs, minus the repeats at
    # the same point that cannot change the cache (fp = func(p) at minimize_nrc.h:634, fx = func(bx) at :315)
    assert (r["status"] != 6).all() and (r["status"] != 7).all() and (r2["status"] < 6).all()


def test_strict_camera_wave_from_file_state(ladybug, oracle_mod):
    """The 49 camera components from the file's state (cameras first), 8 CPU threads on 8 oracle replicas."""
    spec, pts, cams = ladybug
    x0 = spec["x0"]
    ctx, orc = _pair(spec, oracle_mod, x0)
    reps = [orc] + [oracle_mod.OracleFunction.from_spec(spec, "devtrig") for _ in range(7)]
    for rep in reps[1:]:
        rep.set_x(x0)
    r = ctx.solve_cgd(cams, x0[cams.vids], 25, 3e-8)
    o = orc.solve_cgd_batch(cams.var_off, cams.vids, cams.fac_off, cams.fids, x0[cams.vids], 25, 3e-8, replicas=reps)
    _assert_identical(r, o, "camera wave")


@pytest.mark.parametrize("shape", ["subtrees", "chain", "tree", "odd_arity"])
def test_strict_sinusoid_is_bit_identical(built_lib, oracle_mod, shape):
    """NonlinearProductFactor graphs: sibling subtrees of a sinusoid tree, BASELINE config 2 (the d=1000 chain and the
    default-shaped tree) as ONE subspace problem over every variable and factor."""
    from rdis_b200 import problems as P
    if shape == "subtrees":
        tree, lv = P.sinusoid(10, 2, 4), 4
    elif shape == "chain":
        tree, lv = P.sinusoid(999, 1, 3), 0
    elif shape == "tree":
        tree, lv = P.sinusoid(6, 3, 4), 0
    else:
        tree, lv = P.sinusoid(5, 3, 5, odd=True), 2
    xt = P.random_start(tree, 5)
    ctx, orc = _pair(tree, oracle_mod, xt)
    ps = P.sinusoid_subtree_problems(tree, lv) if lv else P.full_problem(tree)
    r = ctx.solve_cgd(ps, xt[ps.vids], 25, 3e-8)
    o = orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, xt[ps.vids], 25, 3e-8)
    _assert_identical(r, o, shape)


def test_strict_change_filter_and_constants(built_lib, oracle_mod):
    """The cache semantics across calls: a state change below 1e-12 through rdisgpu_set_x does not invalidate the
    cached factor values (src/Variable.cpp:69-73), a change above it does; assigned-constant factors contribute their
    constant and keep their gradient (src/Factor.cpp:110-119); clamping into the domain on entry."""
    from rdis_b200 import problems as P
    spec = P.ba_synthetic(ncams=5, npts=60, nobs=260, seed=3)
    x0 = spec["x0"]
    pts, cams = P.ba_point_problems(spec), P.ba_camera_problems(spec)
    ctx, orc = _pair(spec, oracle_mod, x0)
    rng = np.random.default_rng(1)
    fid = rng.permutation(spec["F"])[:40].astype(np.int64)
    val = rng.normal(0, 3, 40)
    on = np.ones(40, np.uint8)
    ctx.set_factor_const(fid, val, on)
    orc.set_factor_const(fid, val, on)
    for wave, ps in enumerate((pts, cams, pts)):
        x = ctx.get_x()
        if wave == 1:  # nudge every variable by less than the filter's tolerance: caches stay valid on both sides
            x = x + rng.uniform(-4e-13, 4e-13, x.size)
            ctx.set_x(x)
            orc.set_x(x)
        if wave == 2:  # release half of the constants, move some variables visibly, start outside the domain
            ctx.set_factor_const(fid[:20], val[:20], np.zeros(20, np.uint8))
            orc.set_factor_const(fid[:20], val[:20], np.zeros(20, np.uint8))
            x[::7] += 1e-9
            ctx.set_x(x)
            orc.set_x(x)
        xs = x[ps.vids].copy()
        if wave == 2:
            xs[::5] = spec["ub"][ps.vids][::5] + 1.0
        r = ctx.solve_cgd(ps, xs, 25, 3e-8)
        o = orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, xs, 25, 3e-8)
        _assert_identical(r, o, "wave %d" % wave)
        assert np.array_equal(_bits(ctx.get_x()), _bits(orc.get_x()))


def test_strict_top_level_block(ladybug, oracle_mod):
    """BASELINE config 3 shape (i): a top-level block of the real ladybug graph — the first 8 cameras and every point
    only they observe... restricted to a size the CPU oracle finishes in seconds: 5 cameras + 300 points chosen
    together (945 variables), all factors whose variables are all assigned."""
    spec, pts, cams = ladybug
    from rdis_b200.problems import ProblemSet
    x0 = spec["x0"]
    ncams = spec["ncams"]
    cam_sel = np.arange(5)
    pt_sel = np.unique(spec["pt"][np.isin(spec["cam"], cam_sel)])[:300]
    vids = np.concatenate([np.arange(9 * c, 9 * c + 9) for c in cam_sel] + [9 * ncams + 3 * pt_sel[:, None] + np.arange(3)[None, :]], axis=None).astype(np.int32)
    vids = np.sort(vids)
    fids = np.nonzero(np.isin(spec["cam"], cam_sel) | np.isin(spec["pt"], pt_sel))[0].astype(np.int64)
    ps = ProblemSet([0, len(vids)], vids, [0, len(fids)], fids)
    ctx, orc = _pair(spec, oracle_mod, x0)
    r = ctx.solve_cgd(ps, x0[vids], 25, 3e-8)
    o = orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[vids], 25, 3e-8)
    _assert_identical(r, o, "top-level block")
