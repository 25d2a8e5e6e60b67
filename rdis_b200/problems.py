"""Synthetic factor graphs of the reference's shapes, as flat arrays in the C-ABI's input format.

    sinusoid(...)        makeHighDimSinusoid, src/OptimizableFunctionGenerator.cpp:660-760
                         (vectorised; tests check it against the oracle's line-by-line restatement)
    ba_synthetic(...)    a bundle-adjustment problem with the shape of
                         data/ladybug-problem-49-7776-pre.txt (49 cameras, 7776 points, 31843
                         observations, 2..29 observations per point)
    ba_domains(...)      BundleAdjustmentFunction::setDomain, src/bundleadjust/BundleAdjustmentFunction.cpp:402-477
and the sibling-component problem sets the recursive decomposer produces on them
(Component::createChildren, src/Component.cpp:508-549; gdfs builder, src/RDISOptimizer.cpp:1049-1059).
Host-side numpy only; nothing here evaluates a factor, and the module has no import inside the package, so
bench.py's reference arm can load it by path without mapping librdis_b200.so.
"""
import os

import numpy as np


class ProblemSet:
    """Host-side description of a batch of subspace problems (CSR style)."""

    def __init__(self, var_off, vids, fac_off, fids):
        self.var_off = np.ascontiguousarray(var_off, dtype=np.int64)
        self.vids = np.ascontiguousarray(vids, dtype=np.int32)
        self.fac_off = np.ascontiguousarray(fac_off, dtype=np.int64)
        self.fids = np.ascontiguousarray(fids, dtype=np.int64)
        self.n = len(self.var_off) - 1
        assert len(self.fac_off) == self.n + 1

    @classmethod
    def from_lists(cls, problems):
        """problems: iterable of (vids, fids)."""
        vo, fo, vs, fs = [0], [0], [], []
        for v, f in problems:
            vs.append(np.asarray(v, np.int32)); fs.append(np.asarray(f, np.int64))
            vo.append(vo[-1] + len(v)); fo.append(fo[-1] + len(f))
        return cls(vo, np.concatenate(vs) if vs else np.zeros(0, np.int32), fo,
                   np.concatenate(fs) if fs else np.zeros(0, np.int64))

    def subset(self, idx):
        return ProblemSet.from_lists([(self.vids[self.var_off[i]:self.var_off[i + 1]],
                                       self.fids[self.fac_off[i]:self.fac_off[i + 1]]) for i in idx])


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# ------------------------------------------------------------------------------------------
# sinusoid tree (NonlinearProductFactor)
# ------------------------------------------------------------------------------------------
def sinusoid(height, branches, max_arity, odd=False):
    h, k = int(height), int(branches)
    twopi = 2.000001 * 3.141592653
    bound = float("%g" % (10 * twopi))  # boost::format default precision, Generator.cpp:669-671
    max_arity = min(int(max_arity), h + 1)
    nvars = h + 1 if k == 1 else (k ** (h + 1) - 1) // (k - 1)
    vid_all = np.arange(nvars, dtype=np.int64)
    if k == 1:
        depth = vid_all.copy()
    else:
        # depth d holds BFS indices [(k^d-1)/(k-1), (k^(d+1)-1)/(k-1))
        starts = np.array([(k ** d - 1) // (k - 1) for d in range(h + 2)], dtype=np.int64)
        depth = np.searchsorted(starts, vid_all, side="right") - 1
    rows_vid, rows_len, coeffs, sines = [], [], [], []
    for ar in range(1, max_arity + 1):
        if ar > 1 and (ar & 1) and not odd:
            continue
        v = vid_all[::-1]
        v = v[depth[v] + 1 >= ar]  # needs ar-1 ancestors (Generator.cpp:712)
        chain = np.empty((len(v), ar), dtype=np.int64)
        cur = v.copy()
        for c in range(ar):  # leaf-to-root walk, stored root-most first (:719-737)
            chain[:, ar - 1 - c] = cur
            cur = (cur - 1) // k
        rows_vid.append(chain.reshape(-1))
        rows_len.append(np.full(len(v), ar, dtype=np.int64))
        coeffs.append(np.full(len(v), 12.0 if ar > 1 else 0.6))
        sines.append(np.full(len(v) * ar, 1 if ar > 1 else 0, dtype=np.uint8))
    n_tree_edges = int(sum(len(r) for r in rows_vid))
    # 0.1 * x^2 per variable (:743-747)
    rows_vid.append(vid_all)
    rows_len.append(np.ones(nvars, dtype=np.int64))
    coeffs.append(np.full(nvars, 0.1))
    sines.append(np.zeros(nvars, dtype=np.uint8))
    vid = np.concatenate(rows_vid).astype(np.int32)
    lens = np.concatenate(rows_len)
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    expo = np.ones(len(vid))
    expo[n_tree_edges:] = 2.0
    return {
        "kind": "nlpf", "V": nvars, "F": len(lens),
        "lb": np.full(nvars, -bound), "ub": np.full(nvars, bound),
        "samp_lo": np.full(nvars, -twopi), "samp_hi": np.full(nvars, twopi),
        "rowptr": rowptr, "vid": vid, "expo": expo, "konst": np.zeros(len(vid)),
        "sine": np.concatenate(sines), "coeff": np.concatenate(coeffs),
        "tree": (h, k),
    }


def random_start(spec, seed):
    """Uniform draw over each variable's sampling interval (what the CLIs do for their initial
    states, src/optimize_sinusoid.cpp:154-165 / src/bundleadjust/optBA.cpp:198-205; our own
    stream — Boost.Random's distribution is un-vendored)."""
    rng = np.random.default_rng(seed)
    return rng.uniform(spec["samp_lo"], spec["samp_hi"])


def nlpf_var_incidence(spec):
    """variable -> factor ids (ascending) as CSR."""
    lens = np.diff(spec["rowptr"])
    efac = np.repeat(np.arange(spec["F"], dtype=np.int64), lens)
    order = np.argsort(spec["vid"], kind="stable")
    vrow = np.concatenate([[0], np.cumsum(np.bincount(spec["vid"], minlength=spec["V"]))])
    return vrow, efac[order]


def sinusoid_subtree_problems(spec, assigned_levels):
    """Sibling components after assigning the top `assigned_levels` levels of the tree: every
    subtree rooted at depth `assigned_levels` is one component; its factors are those whose
    variables are all assigned-or-inside (they then touch only that subtree + fixed ancestors)."""
    h, k = spec["tree"]
    V = spec["V"]
    vid_all = np.arange(V, dtype=np.int64)
    if k == 1:
        depth = vid_all.copy()
    else:
        starts = np.array([(k ** d - 1) // (k - 1) for d in range(h + 2)], dtype=np.int64)
        depth = np.searchsorted(starts, vid_all, side="right") - 1
    # component label of a variable = its ancestor at depth `assigned_levels` (or -1 if assigned)
    label = vid_all.copy()
    for _ in range(h + 1):
        up = depth[label] > assigned_levels
        if not up.any():
            break
        label[up] = (label[up] - 1) // k
    label[depth < assigned_levels] = -1
    # a factor's component = label of its deepest (= last) variable; all its other variables are ancestors
    last_edge = spec["rowptr"][1:] - 1
    flabel = label[spec["vid"][last_edge]]
    probs = []
    roots = vid_all[depth == assigned_levels]
    vorder = np.argsort(label, kind="stable")
    forder = np.argsort(flabel, kind="stable")
    vl, fl = label[vorder], flabel[forder]
    for r in roots:
        vs = vorder[np.searchsorted(vl, r, "left"):np.searchsorted(vl, r, "right")]
        fs = forder[np.searchsorted(fl, r, "left"):np.searchsorted(fl, r, "right")]
        probs.append((np.sort(vs).astype(np.int32), np.sort(fs).astype(np.int64)))
    return ProblemSet.from_lists(probs)


# ------------------------------------------------------------------------------------------
# bundle adjustment
# ------------------------------------------------------------------------------------------
def ba_domains(x0, ncams):
    """Domains and sampling intervals of BundleAdjustmentFunction::setDomain (no-rounding intervals)."""
    V = len(x0)
    slot = np.where(np.arange(V) < 9 * ncams, np.arange(V) % 9, 9 + (np.arange(V) - 9 * ncams) % 3)
    dsf = 1000.0
    slo = np.empty(V); shi = np.empty(V); dlo = np.empty(V); dhi = np.empty(V)
    rot = slot <= 2
    slo[rot], shi[rot] = -np.pi, np.pi
    dlo[rot], dhi[rot] = -np.pi * dsf, np.pi * dsf
    pos = ((slot >= 3) & (slot <= 5)) | (slot >= 9)
    slo[pos], shi[pos] = x0[pos] - 1e2, x0[pos] + 1e2
    a, b = slo[pos] * dsf, shi[pos] * dsf
    dlo[pos], dhi[pos] = np.minimum(a, b), np.maximum(a, b)
    foc = slot == 6
    slo[foc], shi[foc] = x0[foc] - 1e2, x0[foc] + 1e2
    dlo[foc] = np.maximum(np.minimum(slo[foc], slo[foc] * dsf), 0.0)
    dhi[foc] = shi[foc] * dsf
    k1 = slot == 7
    slo[k1], shi[k1] = x0[k1] - 1e-4, x0[k1] + 1e-4
    dlo[k1], dhi[k1] = -1e-1, 1e-1
    k2 = slot == 8
    slo[k2], shi[k2] = x0[k2] - 1e-6, x0[k2] + 1e-6
    dlo[k2], dhi[k2] = -1e-3, 1e-3
    return np.minimum(dlo, slo), np.maximum(dhi, shi), slo, shi


def _project(cams, pts, cam_idx, pt_idx):
    """Reprojection in numpy, used ONLY to synthesise observations for generated problems."""
    r = cams[cam_idx, 0:3]
    q = pts[pt_idx]
    th = np.linalg.norm(r, axis=1, keepdims=True)
    a = r / np.where(th == 0, 1.0, th)
    c, s = np.cos(th), np.sin(th)
    P = q * c + np.cross(a, q) * s + a * (1 - c) * np.sum(a * q, axis=1, keepdims=True)
    P = P + cams[cam_idx, 3:6]
    pp = -P[:, :2] / P[:, 2:3]
    r2 = np.sum(pp * pp, axis=1, keepdims=True)
    d = 1 + r2 * (cams[cam_idx, 7:8] + cams[cam_idx, 8:9] * r2)
    return cams[cam_idx, 6:7] * d * pp


def ba_synthetic(ncams=49, npts=7776, nobs=31843, seed=20260417, noise_px=0.5, perturb=1.0):
    """Ladybug-shaped synthetic BA problem.  Observations per point: min 2, median 3, max 29,
    exactly `nobs` in total; every camera sees several hundred points.  x0 is a perturbed copy
    of the generating state so that solves have real work to do."""
    rng = np.random.default_rng(seed)
    max_deg = min(29, ncams)
    if ncams < 2 or not (2 * npts <= nobs <= max_deg * npts):
        raise ValueError("ba_synthetic: nobs must lie in [2 * npts, min(29, ncams) * npts] (every point is seen by 2..%d cameras)" % max_deg)
    deg = 2 + np.minimum(rng.geometric(0.42, size=npts) - 1, max_deg - 2)
    # hit nobs exactly
    diff = int(nobs - deg.sum())
    while diff != 0:
        i = rng.integers(0, npts, size=abs(diff))
        if diff > 0:
            ok = deg[i] < max_deg
            np.add.at(deg, i[ok], 1)
        else:
            ok = deg[i] > 2
            np.subtract.at(deg, i[ok], 1)
        deg = np.clip(deg, 2, max_deg)
        diff = int(nobs - deg.sum())
    cam_w = rng.uniform(0.6, 1.5, size=ncams)
    cam_w /= cam_w.sum()
    cam_idx = np.concatenate([rng.choice(ncams, size=d, replace=False, p=cam_w) for d in deg]).astype(np.int32)
    pt_idx = np.repeat(np.arange(npts, dtype=np.int32), deg)
    # BAL files list observations point-major with ascending camera id
    order = np.lexsort((cam_idx, pt_idx))
    cam_idx, pt_idx = cam_idx[order], pt_idx[order]

    cams = np.empty((ncams, 9))
    cams[:, 0:3] = rng.normal(0, 0.15, size=(ncams, 3))
    cams[:, 3:6] = rng.normal(0, 0.8, size=(ncams, 3))
    cams[:, 6] = rng.uniform(380, 420, size=ncams)
    cams[:, 7] = rng.normal(-3e-7, 5e-8, size=ncams)
    cams[:, 8] = rng.normal(6e-13, 1e-13, size=ncams)
    pts = np.empty((npts, 3))
    pts[:, 0:2] = rng.normal(0, 4.0, size=(npts, 2))
    pts[:, 2] = rng.uniform(-30, -12, size=npts)  # in front of the cameras (BAL: camera looks down -z)
    obs = _project(cams, pts, cam_idx, pt_idx) + rng.normal(0, noise_px, size=(len(cam_idx), 2))

    x_true = np.concatenate([cams.reshape(-1), pts.reshape(-1)])
    cam_sig = np.tile(np.array([2e-3, 2e-3, 2e-3, 2e-2, 2e-2, 2e-2, 0.5, 1e-8, 1e-14]), ncams)
    x0 = x_true + perturb * np.concatenate([rng.normal(0, 1, 9 * ncams) * cam_sig, rng.normal(0, 0.1, 3 * npts)])
    lb, ub, slo, shi = ba_domains(x0, ncams)
    return {"kind": "ba", "V": len(x0), "F": len(cam_idx), "ncams": ncams, "npts": npts,
            "cam": cam_idx, "pt": pt_idx, "obs": obs, "lb": lb, "ub": ub, "samp_lo": slo, "samp_hi": shi, "x0": x0}


def ba_replicate(spec, K):
    """K independent copies of a bundle-adjustment problem in one graph (the K restarts of `optBA --nsamples`,
    src/bundleadjust/optBA.cpp:184-211, as one function): cameras of copy k are cameras k*ncams.., points likewise."""
    nc, npnt = spec["ncams"], spec["npts"]
    cam = np.concatenate([spec["cam"] + k * nc for k in range(K)]).astype(np.int32)
    pt = np.concatenate([spec["pt"] + k * npnt for k in range(K)]).astype(np.int32)
    obs = np.concatenate([np.asarray(spec["obs"]).reshape(-1, 2)] * K)
    x0 = np.concatenate([np.tile(spec["x0"][:9 * nc], K), np.tile(spec["x0"][9 * nc:], K)])
    lb, ub, slo, shi = ba_domains(x0, nc * K)
    return {"kind": "ba", "V": len(x0), "F": spec["F"] * K, "ncams": nc * K, "npts": npnt * K, "cam": cam, "pt": pt, "obs": obs,
            "lb": lb, "ub": ub, "samp_lo": slo, "samp_hi": shi, "x0": x0}


def ba_replicate_points(spec, K):
    """The same cameras observing K copies of the point cloud: ncams cameras, K*npts points, K*F observations
    (a long sequence from few cameras; used to put the BA sweep into its streaming regime)."""
    nc, npnt = spec["ncams"], spec["npts"]
    cam = np.tile(spec["cam"], K).astype(np.int32)
    pt = np.concatenate([spec["pt"] + k * npnt for k in range(K)]).astype(np.int32)
    obs = np.concatenate([np.asarray(spec["obs"]).reshape(-1, 2)] * K)
    x0 = np.concatenate([spec["x0"][:9 * nc], np.tile(spec["x0"][9 * nc:], K)])
    lb, ub, slo, shi = ba_domains(x0, nc)
    return {"kind": "ba", "V": len(x0), "F": spec["F"] * K, "ncams": nc, "npts": npnt * K, "cam": cam, "pt": pt, "obs": obs,
            "lb": lb, "ub": ub, "samp_lo": slo, "samp_hi": shi, "x0": x0}


def load_golden_ba(name="ladybug_49_7776.npz"):
    """The reference's own data/ladybug-problem-49-7776-pre.txt as parsed by the oracle's BAL
    loader (tests/golden/make_golden.py wrote it)."""
    z = np.load(os.path.join(GOLDEN_DIR, name))
    ncams, npts = int(z["ncams"]), int(z["npts"])
    x0 = z["x0"]
    lb, ub, slo, shi = ba_domains(x0, ncams)
    return {"kind": "ba", "V": len(x0), "F": len(z["cam"]), "ncams": ncams, "npts": npts,
            "cam": z["cam"].astype(np.int32), "pt": z["pt"].astype(np.int32), "obs": z["obs"].astype(np.float64),
            "lb": lb, "ub": ub, "samp_lo": slo, "samp_hi": shi, "x0": x0}


def _group_by(keys, n):
    order = np.argsort(keys, kind="stable")
    row = np.concatenate([[0], np.cumsum(np.bincount(keys, minlength=n))])
    return row, order


def ba_point_problems(spec):
    """Cameras assigned: every point is its own 3-variable component whose factors are its
    observations (ascending factor id)."""
    nc, npnt = spec["ncams"], spec["npts"]
    row, order = _group_by(spec["pt"], npnt)
    vids = (9 * nc + np.arange(3 * npnt)).astype(np.int32)
    return ProblemSet(np.arange(0, 3 * npnt + 1, 3), vids, row, order.astype(np.int64))


def ba_camera_problems(spec):
    """Points assigned: every camera is its own 9-variable component."""
    nc = spec["ncams"]
    row, order = _group_by(spec["cam"], nc)
    return ProblemSet(np.arange(0, 9 * nc + 1, 9), np.arange(9 * nc, dtype=np.int32), row, order.astype(np.int64))


def full_problem(spec):
    return ProblemSet([0, spec["V"]], np.arange(spec["V"], dtype=np.int32), [0, spec["F"]],
                      np.arange(spec["F"], dtype=np.int64))


def shard_problems(ps, rank, world):
    from .shard import shard_problems as _impl
    return _impl(ps, rank, world)
