"""The C++ host layer (rdis_b200/host): the reference's plugin surface kept on top of the C-ABI.

CPU: ComponentBatcher::createChildren (component membership — integer bookkeeping, must be exact)
     against an independent connected-components labelling (scipy) of the same bipartite graph.
GPU: one sibling wave through CudaSubspaceOptimizer::optimizeBatch and through the reference-style
     loop of optimize() calls, against the ctypes path (bit-exact: same library) and the oracle (1e-6).
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "native", "host_driver")


@pytest.fixture(scope="module")
def driver(built_lib):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "rdis_b200", "host"), "all"], stdout=subprocess.DEVNULL)
    assert os.path.exists(DRIVER)
    return DRIVER


def write_problem(path, spec, x0, assigned):
    with open(path, "wb") as fh:
        nlpf = spec["kind"] == "nlpf"
        E = len(spec["vid"]) if nlpf else 0
        np.array([0 if nlpf else 1, spec["V"], spec["F"], E, spec.get("ncams", 0), spec.get("npts", 0)], np.int64).tofile(fh)
        np.asarray(spec["lb"], np.float64).tofile(fh)
        np.asarray(spec["ub"], np.float64).tofile(fh)
        if nlpf:
            np.asarray(spec["rowptr"], np.int64).tofile(fh)
            np.asarray(spec["vid"], np.int32).tofile(fh)
            np.asarray(spec["expo"], np.float64).tofile(fh)
            np.asarray(spec["konst"], np.float64).tofile(fh)
            np.asarray(spec["sine"], np.uint8).tofile(fh)
            np.asarray(spec["coeff"], np.float64).tofile(fh)
        else:
            np.asarray(spec["cam"], np.int32).tofile(fh)
            np.asarray(spec["pt"], np.int32).tofile(fh)
            np.asarray(spec["obs"], np.float64).reshape(-1).tofile(fh)
        np.asarray(x0, np.float64).tofile(fh)
        np.asarray(assigned, np.uint8).tofile(fh)


def factor_vars(spec):
    """factor id -> array of variable ids"""
    if spec["kind"] == "nlpf":
        rp = spec["rowptr"]
        return [spec["vid"][rp[j]:rp[j + 1]].astype(np.int64) for j in range(spec["F"])]
    nc = spec["ncams"]
    return [np.concatenate([9 * int(c) + np.arange(9), 9 * nc + 3 * int(p) + np.arange(3)]) for c, p in zip(spec["cam"], spec["pt"])]


def reference_children(spec, assigned):
    """Independent restatement of Component::createChildren's OUTPUT (src/Component.cpp:508-549, :50-80,
    :603-608): connected components of the var/factor graph restricted to unassigned variables."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    V, F = spec["V"], spec["F"]
    fv = factor_vars(spec)
    rows, cols = [], []
    for j, vs in enumerate(fv):
        for v in vs:
            if not assigned[v]:
                rows.append(int(v)); cols.append(V + j)
    g = coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(V + F, V + F))
    _, lab = connected_components(g, directed=False)
    kids = {}
    for v in range(V):
        if not assigned[v]:
            kids.setdefault(lab[v], ([], []))[0].append(v)
    for j in range(F):
        if lab[V + j] in kids and any(not assigned[v] for v in fv[j]):
            kids[lab[V + j]][1].append(j)
    out = [(sorted(vs), sorted(fs)) for vs, fs in kids.values()]
    out.sort(key=lambda c: (len(c[0]), c[0][0]))
    return out


def parse_children(text):
    lines = text.strip().splitlines()
    n = int(lines[0].split()[1])
    kids = []
    for ln in lines[1:1 + n]:
        head, vs, fs = ln.split("|")
        kids.append(([int(t) for t in vs.split()], [int(t) for t in fs.split()]))
    return kids


def test_children_membership_is_exact_cpu(driver, tmp_path):
    from rdis_b200 import problems as P
    cases = []
    spec = P.ba_synthetic(ncams=5, npts=40, nobs=150, seed=3)
    a = np.zeros(spec["V"], np.uint8); a[:9 * 5] = 1                      # cameras assigned: one child per point
    cases.append((spec, a))
    a = np.zeros(spec["V"], np.uint8); a[9 * 5:] = 1                      # points assigned: one child per camera
    cases.append((spec, a))
    a = np.zeros(spec["V"], np.uint8); a[:9 * 2] = 1; a[9 * 5:9 * 5 + 30] = 1  # a mixed cut
    cases.append((spec, a))
    sp2 = P.sinusoid(6, 2, 4)
    a = np.zeros(sp2["V"], np.uint8); a[:7] = 1                           # top three tree levels assigned: 8 subtrees
    cases.append((sp2, a))
    a = np.zeros(sp2["V"], np.uint8); a[::3] = 1
    cases.append((sp2, a))
    for i, (sp, a) in enumerate(cases):
        path = str(tmp_path / ("p%d.bin" % i))
        write_problem(path, sp, np.zeros(sp["V"]), a)
        out = subprocess.run([driver, "children", path], capture_output=True, text=True, check=True).stdout
        got = parse_children(out)
        want = reference_children(sp, a)
        assert got == want, "case %d: component membership differs" % i
    # the shapes the bench relies on
    sp, a = cases[0]
    kids = reference_children(sp, a)
    assert len(kids) == 40 and all(len(v) == 3 for v, _ in kids)


def _parse_wave(text):
    lines = text.strip().splitlines()
    n = int(lines[0].split()[1]); total = float(lines[0].split()[3])
    probs = []
    for ln in lines[1:1 + n]:
        head, xs = ln.split("|")
        fval, dfv, nv, nf = head.split()
        x = {int(t.split(":")[0]): float(t.split(":")[1]) for t in xs.split()}
        probs.append((float(fval), float(dfv), int(nv), int(nf), x))
    rest = {ln.split()[0]: [float(t) for t in ln.split()[1:]] for ln in lines[1 + n:]}
    return total, probs, rest


@pytest.mark.gpu
def test_host_adapter_wave_matches_ctypes_path_and_oracle(driver, oracle_mod, tmp_path):
    from rdis_b200 import Context, problems as P
    spec = P.ba_synthetic(ncams=6, npts=120, nobs=520, seed=9)
    x0 = spec["x0"]
    assigned = np.zeros(spec["V"], np.uint8); assigned[:9 * 6] = 1   # cameras assigned, points are the sibling wave
    path = str(tmp_path / "wave.bin")
    write_problem(path, spec, x0, assigned)
    out_w = subprocess.run([driver, "wave", path, "25"], capture_output=True, text=True, check=True).stdout
    out_s = subprocess.run([driver, "single", path, "25"], capture_output=True, text=True, check=True).stdout
    tot_w, pw, rest_w = _parse_wave(out_w)
    tot_s, ps_, rest_s = _parse_wave(out_s)
    assert "POSTCONDITION" not in out_w and "POSTCONDITION" not in out_s
    # batch == reference-style loop of optimize() calls, to the bit
    assert pw == ps_ and rest_w == rest_s
    # == the ctypes path on the same problems (children come out in (size, first id) order = point order here)
    ps = P.ba_point_problems(spec)
    ctx = Context.from_spec(spec); ctx.set_x(x0)
    r = ctx.solve_cgd(ps, x0[ps.vids], 25, 3e-8)
    assert len(pw) == ps.n
    for k, (fval, dfv, nv, nf, x) in enumerate(pw):
        vids = ps.vids[ps.var_off[k]:ps.var_off[k + 1]]
        assert sorted(x) == vids.tolist()
        assert fval == r["f_end"][k] and dfv == r["f_end"][k] - r["f_init"][k]
        assert [x[int(v)] for v in vids] == r["x"][ps.var_off[k]:ps.var_off[k + 1]].tolist()
    assert abs(rest_w["eval_all"][0] - ctx.eval()) <= 1e-12 * abs(ctx.eval())
    g = ctx.grad(vid=ps.vids[:3])
    assert np.allclose(rest_w["grad0"], g, rtol=1e-13, atol=0)
    # vs the oracle
    orc = oracle_mod.OracleFunction.from_spec(spec); orc.set_x(x0)
    o = orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[ps.vids], 25, 3e-8)
    rel = np.abs(np.array([p[0] for p in pw]) - o["f_end"]) / np.maximum(np.abs(o["f_end"]), 1e-12)
    assert rel.max() <= 1e-6
    # the Levenberg-Marquardt adapter (CudaLMSubspaceOptimizer) through the same surface
    out_l = subprocess.run([driver, "wave", path, "25", "lm"], capture_output=True, text=True, check=True).stdout
    tot_l, pl, _ = _parse_wave(out_l)
    orc.set_x(x0)
    ol = orc.solve_lm_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[ps.vids], 25, 3e-8)
    rel = np.abs(np.array([p[0] for p in pl]) - ol["f_end"]) / np.maximum(np.abs(ol["f_end"]), 1e-12)
    assert rel.max() <= 1e-6


@pytest.mark.gpu
def test_host_adapter_nlpf_subtrees(driver, oracle_mod, tmp_path):
    from rdis_b200 import problems as P
    spec = P.sinusoid(7, 2, 4)
    x0 = P.random_start(spec, 4)
    assigned = np.zeros(spec["V"], np.uint8); assigned[:7] = 1
    path = str(tmp_path / "nlpf.bin")
    write_problem(path, spec, x0, assigned)
    out = subprocess.run([driver, "wave", path, "25"], capture_output=True, text=True, check=True).stdout
    tot, pw, rest = _parse_wave(out)
    ps = P.sinusoid_subtree_problems(spec, 3)
    assert len(pw) == ps.n == 8
    orc = oracle_mod.OracleFunction.from_spec(spec); orc.set_x(x0)
    o = orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[ps.vids], 25, 3e-8)
    got = np.array([p[0] for p in pw])
    rel = np.abs(got - o["f_end"]) / np.maximum(np.abs(o["f_end"]), 1e-12)
    print("nlpf subtrees through the C++ adapter: worst rel diff", rel.max())
    assert rel.max() <= 1e-6
