"""ctypes binding of the C-ABI declared in include/rdis_gpu.h (librdis_b200.so).

This is plumbing for tests and bench.py: every call goes straight to the hand-written
CUDA library.  There is no CPU fallback — if the shared library is missing the import
fails, and if no sm_100 GPU is usable `Context()` raises.
"""
import ctypes as C
import os
import weakref

import numpy as np

from .problems import ProblemSet  # noqa: F401  (numpy only; re-exported)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RDIS_B200_LIB") or os.path.join(_HERE, "librdis_b200.so")  # override: kernel experiments only

DONE_NAMES = ("ftol", "gtol", "gg_zero", "maxiters", "dbrent_itmax", "empty", "nonfinite", "bracket_cap")

# every symbol include/rdis_gpu.h declares (tests check the library exports all of them)
EXPORTS = (
    "rdisgpu_create", "rdisgpu_destroy", "rdisgpu_last_error", "rdisgpu_set_stream", "rdisgpu_synchronize", "rdisgpu_set_option",
    "rdisgpu_set_vars", "rdisgpu_add_nlpf", "rdisgpu_add_ba", "rdisgpu_finalize",
    "rdisgpu_set_x", "rdisgpu_get_x", "rdisgpu_set_factor_const",
    "rdisgpu_eval", "rdisgpu_grad", "rdisgpu_eval_device", "rdisgpu_grad_device", "rdisgpu_factor_grad", "rdisgpu_factor_rows_device",
    "rdisgpu_solve_cgd", "rdisgpu_solve_cgd_csr", "rdisgpu_solve_lm_csr", "rdisgpu_batch_create", "rdisgpu_batch_create_csr", "rdisgpu_batch_info", "rdisgpu_batch_solve_cgd", "rdisgpu_batch_fetch", "rdisgpu_batch_fetch_csr",
    "rdisgpu_batch_objective_device", "rdisgpu_batch_destroy", "rdisgpu_batch_last_launches", "rdisgpu_batch_resident_info", "rdisgpu_components", "rdisgpu_bounds", "rdisgpu_bounds_lists",
    "rdisgpu_num_vars", "rdisgpu_num_factors", "rdisgpu_device_state", "rdisgpu_launch_count", "rdisgpu_version",
)


class Problem(C.Structure):
    _fields_ = [("nv", C.c_int64), ("vid", C.c_void_p), ("nf", C.c_int64), ("fid", C.c_void_p), ("x0", C.c_void_p)]


class Result(C.Structure):
    _fields_ = [("x", C.c_void_p), ("f_init", C.c_double), ("f_end", C.c_double), ("iters", C.c_int32),
                ("status", C.c_int32), ("n_feval", C.c_int64), ("n_geval", C.c_int64)]


class RdisGpuError(RuntimeError):
    pass


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  rdis_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    sig = {
        "rdisgpu_create": (C.c_int, [C.POINTER(vp), C.c_int]),
        "rdisgpu_destroy": (None, [vp]),
        "rdisgpu_last_error": (C.c_char_p, [vp]),
        "rdisgpu_set_stream": (C.c_int, [vp, vp]),
        "rdisgpu_synchronize": (C.c_int, [vp]),
        "rdisgpu_set_option": (C.c_int, [vp, C.c_char_p, i64]),
        "rdisgpu_set_vars": (C.c_int, [vp, i64, vp, vp]),
        "rdisgpu_add_nlpf": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp]),
        "rdisgpu_add_ba": (C.c_int, [vp, i64, vp, vp, vp, i32, i32]),
        "rdisgpu_finalize": (C.c_int, [vp]),
        "rdisgpu_set_x": (C.c_int, [vp, i64, vp, vp]),
        "rdisgpu_get_x": (C.c_int, [vp, i64, vp, vp]),
        "rdisgpu_set_factor_const": (C.c_int, [vp, i64, vp, vp, vp]),
        "rdisgpu_eval": (C.c_int, [vp, i64, vp, C.POINTER(dbl), vp]),
        "rdisgpu_grad": (C.c_int, [vp, i64, vp, i64, vp, vp]),
        "rdisgpu_eval_device": (C.c_int, [vp, i64, vp, vp, vp]),
        "rdisgpu_grad_device": (C.c_int, [vp, i64, vp, i64, vp, vp]),
        "rdisgpu_factor_grad": (C.c_int, [vp, i64, vp, i32, vp]),
        "rdisgpu_factor_rows_device": (C.c_int, [vp, vp, vp, vp]),
        "rdisgpu_solve_cgd": (C.c_int, [vp, C.POINTER(Problem), i64, C.c_int, dbl, C.POINTER(Result)]),
        "rdisgpu_batch_create": (C.c_int, [vp, C.POINTER(Problem), i64, C.POINTER(vp)]),
        "rdisgpu_batch_create_csr": (C.c_int, [vp, i64, vp, vp, vp, vp, C.POINTER(vp)]),
        "rdisgpu_solve_cgd_csr": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, C.c_int, dbl, vp, vp, vp, vp, vp, vp, vp]),
        "rdisgpu_solve_lm_csr": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]),
        "rdisgpu_components": (C.c_int, [vp, vp, vp, vp, C.POINTER(i32), C.POINTER(i32)]),
        "rdisgpu_bounds": (C.c_int, [vp, vp, i64, vp, vp, vp, vp]),
        "rdisgpu_bounds_lists": (C.c_int, [vp, vp, i64, vp, vp, vp]),
        "rdisgpu_batch_info": (C.c_int, [vp, vp]),
        "rdisgpu_batch_resident_info": (C.c_int, [vp, vp]),
        "rdisgpu_batch_solve_cgd": (C.c_int, [vp, vp, C.c_int, dbl]),
        "rdisgpu_batch_fetch": (C.c_int, [vp, C.POINTER(Result), C.POINTER(dbl)]),
        "rdisgpu_batch_fetch_csr": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp]),
        "rdisgpu_batch_objective_device": (C.c_int, [vp, vp]),
        "rdisgpu_batch_destroy": (None, [vp]),
        "rdisgpu_batch_last_launches": (C.c_int, [vp]),
        "rdisgpu_num_vars": (i64, [vp]),
        "rdisgpu_num_factors": (i64, [vp]),
        "rdisgpu_device_state": (C.c_int, [vp, C.POINTER(vp), C.POINTER(i64)]),
        "rdisgpu_launch_count": (i64, [vp]),
        "rdisgpu_version": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def problem_array(ps, x0=None):
    """The rdisgpu_problem array of a ProblemSet (pointers into its arrays: keep `ps` and `x0` alive)."""
    arr = (Problem * ps.n)()
    vbase, fbase = ps.vids.ctypes.data, ps.fids.ctypes.data
    xbase = None if x0 is None else x0.ctypes.data
    for i in range(ps.n):
        arr[i].nv = int(ps.var_off[i + 1] - ps.var_off[i])
        arr[i].nf = int(ps.fac_off[i + 1] - ps.fac_off[i])
        arr[i].vid = vbase + 4 * int(ps.var_off[i])
        arr[i].fid = fbase + 8 * int(ps.fac_off[i])
        arr[i].x0 = None if xbase is None else xbase + 8 * int(ps.var_off[i])
    return arr


class Context:
    """One OptimizableFunction resident on one GPU (rdisgpu_ctx)."""

    def __init__(self, device=0):
        self._lib = lib()
        h = C.c_void_p()
        rc = self._lib.rdisgpu_create(C.byref(h), device)
        if rc != 0:
            raise RdisGpuError(f"rdisgpu_create failed ({rc}): {self._lib.rdisgpu_last_error(None).decode()}")
        self._h = h
        self._keep = []
        self._batches = weakref.WeakSet()   # live Batch objects: closed before the context (they point into it)

    def close(self):
        if getattr(self, "_h", None):
            for b in list(getattr(self, "_batches", ())):
                b.close()
            self._lib.rdisgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RdisGpuError(f"rdisgpu error {rc}: {self._lib.rdisgpu_last_error(self._h).decode()}")

    # ---- definition ------------------------------------------------------------------
    @classmethod
    def from_spec(cls, spec, device=0, stream=None):
        ctx = cls(device)
        if stream is not None:
            ctx.set_stream(stream)
        lb = _arr(spec["lb"], np.float64); ub = _arr(spec["ub"], np.float64)
        ctx._ck(ctx._lib.rdisgpu_set_vars(ctx._h, len(lb), _p(lb), _p(ub)))
        if spec["kind"] == "ba":
            cam = _arr(spec["cam"], np.int32); pt = _arr(spec["pt"], np.int32)
            obs = _arr(spec["obs"], np.float64).reshape(-1)
            ctx._ck(ctx._lib.rdisgpu_add_ba(ctx._h, len(cam), _p(cam), _p(pt), _p(obs), spec["ncams"], spec["npts"]))
        else:
            rp = _arr(spec["rowptr"], np.int64); vid = _arr(spec["vid"], np.int32)
            ex = _arr(spec["expo"], np.float64); ko = _arr(spec["konst"], np.float64)
            si = _arr(spec["sine"], np.uint8); co = _arr(spec["coeff"], np.float64)
            ctx._ck(ctx._lib.rdisgpu_add_nlpf(ctx._h, len(co), _p(rp), _p(vid), _p(ex), _p(ko), _p(si), _p(co)))
        ctx._ck(ctx._lib.rdisgpu_finalize(ctx._h))
        return ctx

    def set_stream(self, cuda_stream_ptr):
        self._ck(self._lib.rdisgpu_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._ck(self._lib.rdisgpu_synchronize(self._h))

    def set_option(self, name, value):
        self._ck(self._lib.rdisgpu_set_option(self._h, name.encode(), int(value)))

    @property
    def V(self):
        return self._lib.rdisgpu_num_vars(self._h)

    @property
    def F(self):
        return self._lib.rdisgpu_num_factors(self._h)

    @property
    def launch_count(self):
        return self._lib.rdisgpu_launch_count(self._h)

    # ---- state -----------------------------------------------------------------------
    def set_x(self, x, vid=None):
        x = _arr(x, np.float64)
        v = None if vid is None else _arr(vid, np.int32)
        self._ck(self._lib.rdisgpu_set_x(self._h, len(x), _p(v), _p(x)))

    def get_x(self, vid=None):
        v = None if vid is None else _arr(vid, np.int32)
        n = self.V if v is None else len(v)
        out = np.empty(n)
        self._ck(self._lib.rdisgpu_get_x(self._h, n, _p(v), _p(out)))
        return out

    def set_x_device(self, x_dev_ptr, n, vid_dev_ptr=None):
        """Asynchronous Variable::assign from device memory (raw pointers)."""
        self._ck(self._lib.rdisgpu_set_x(self._h, n, C.c_void_p(vid_dev_ptr) if vid_dev_ptr else None,
                                         C.c_void_p(x_dev_ptr)))

    def set_factor_const(self, fid, val, on):
        fid = _arr(fid, np.int64); val = _arr(val, np.float64); on = _arr(on, np.uint8)
        self._ck(self._lib.rdisgpu_set_factor_const(self._h, len(fid), _p(fid), _p(val), _p(on)))

    def device_state(self):
        p = C.c_void_p(); n = C.c_int64()
        self._ck(self._lib.rdisgpu_device_state(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    # ---- sweeps ----------------------------------------------------------------------
    def eval(self, fid=None, per_factor=False):
        f = None if fid is None else _arr(fid, np.int64)
        n = self.F if f is None else len(f)
        s = C.c_double(0)
        pf = np.empty(n) if per_factor else None
        self._ck(self._lib.rdisgpu_eval(self._h, n, _p(f), C.byref(s), _p(pf)))
        return (s.value, pf) if per_factor else s.value

    def grad(self, fid=None, vid=None):
        f = None if fid is None else _arr(fid, np.int64)
        v = None if vid is None else _arr(vid, np.int32)
        nv = self.V if v is None else len(v)
        g = np.empty(nv)
        self._ck(self._lib.rdisgpu_grad(self._h, self.F if f is None else len(f), _p(f), nv, _p(v), _p(g)))
        return g

    def eval_device(self, sum_dev_ptr, per_factor_dev_ptr=None, fid_dev_ptr=None, nf=0):
        """Asynchronous residual sweep, raw device pointers (None = all factors / no per-factor output)."""
        self._ck(self._lib.rdisgpu_eval_device(self._h, nf, C.c_void_p(fid_dev_ptr) if fid_dev_ptr else None,
                                               C.c_void_p(sum_dev_ptr) if sum_dev_ptr else None,
                                               C.c_void_p(per_factor_dev_ptr) if per_factor_dev_ptr else None))

    def factor_rows_device(self, per_factor_dev_ptr, rows_dev_ptr, sum_dev_ptr=None):
        """Asynchronous residual + Jacobian-rows sweep over all factors of a BA graph (raw device pointers)."""
        self._ck(self._lib.rdisgpu_factor_rows_device(self._h, C.c_void_p(sum_dev_ptr) if sum_dev_ptr else None,
                                                      C.c_void_p(per_factor_dev_ptr) if per_factor_dev_ptr else None,
                                                      C.c_void_p(rows_dev_ptr)))

    def grad_device(self, g_dev_ptr, vid_dev_ptr=None, nv=0, fid_dev_ptr=None, nf=0):
        """Asynchronous gradient sweep, raw device pointers (None = all variables / all factors)."""
        self._ck(self._lib.rdisgpu_grad_device(self._h, nf, C.c_void_p(fid_dev_ptr) if fid_dev_ptr else None, nv,
                                               C.c_void_p(vid_dev_ptr) if vid_dev_ptr else None, C.c_void_p(g_dev_ptr)))

    def factor_grad(self, fid, arity_max):
        f = _arr(fid, np.int64)
        rows = np.zeros((len(f), arity_max))
        self._ck(self._lib.rdisgpu_factor_grad(self._h, len(f), _p(f), arity_max, _p(rows)))
        return rows

    # ---- interval bounds -------------------------------------------------------------
    def bounds(self, assigned, fids=None):
        """rdisgpu_bounds: (lower[nf], upper[nf], (sum_lower, sum_upper)) of the listed factors (None = all)."""
        a = _arr(assigned, np.uint8)
        assert len(a) == self.V
        f = None if fids is None else _arr(fids, np.int64)
        n = self.F if f is None else len(f)
        lo = np.empty(n); hi = np.empty(n); tot = np.zeros(2)
        self._ck(self._lib.rdisgpu_bounds(self._h, _p(a), n, None if f is None else _p(f), _p(lo), _p(hi), _p(tot)))
        return lo, hi, (float(tot[0]), float(tot[1]))

    def bounds_lists(self, assigned, list_off, fids):
        """rdisgpu_bounds_lists: sums[nlists, 2] = (lower, upper) of every factor list, folded on the device in list order."""
        a = _arr(assigned, np.uint8)
        off = _arr(list_off, np.int64)
        f = _arr(fids, np.int64)
        out = np.zeros((len(off) - 1, 2))
        self._ck(self._lib.rdisgpu_bounds_lists(self._h, _p(a), len(off) - 1, _p(off), _p(f), _p(out)))
        return out

    # ---- component membership ----------------------------------------------------------
    def components(self, assigned):
        """rdisgpu_components: (var_label[V], fac_label[F], n_components, rounds)."""
        a = _arr(assigned, np.uint8)
        assert len(a) == self.V
        vl = np.empty(self.V, np.int32); fl = np.empty(self.F, np.int32)
        n = C.c_int32(0); r = C.c_int32(0)
        self._ck(self._lib.rdisgpu_components(self._h, _p(a), _p(vl), _p(fl), C.byref(n), C.byref(r)))
        return vl, fl, n.value, r.value

    def component_problems(self, assigned):
        """Sibling components as a ProblemSet, ordered like Component::createChildren's children (fewest variables
        first, ties by smallest variable id), variables and factors of a component ascending."""
        vl, fl, n, _ = self.components(assigned)
        vs = np.nonzero(vl >= 0)[0]
        fs = np.nonzero(fl >= 0)[0]
        vorder = vs[np.argsort(vl[vs], kind="stable")]
        forder = fs[np.argsort(fl[fs], kind="stable")]
        labels, vcount = np.unique(vl[vs], return_counts=True)
        fcount = np.zeros(len(labels), np.int64)
        flab, fc = np.unique(fl[fs], return_counts=True)
        fcount[np.searchsorted(labels, flab)] = fc
        voff = np.concatenate([[0], np.cumsum(vcount)]); foff = np.concatenate([[0], np.cumsum(fcount)])
        order = np.lexsort((labels, vcount))   # by size, then by smallest variable id (= the label)
        probs = [(vorder[voff[k]:voff[k + 1]].astype(np.int32), forder[foff[k]:foff[k + 1]].astype(np.int64)) for k in order]
        return ProblemSet.from_lists(probs)

    # ---- solves ----------------------------------------------------------------------
    def solve_cgd(self, problems, x0=None, maxiters=25, ftol=3e-8):
        """rdisgpu_solve_cgd_csr: host buffers in, host buffers out, one C call for the whole batch.
        problems: ProblemSet; x0: concatenated start values or None (use device state).
        Returns dict(x, f_init, f_end, iters, status, n_feval, n_geval)."""
        ps = problems
        x0a = None if x0 is None else _arr(x0, np.float64)
        n = ps.n
        out = {"x": np.empty(len(ps.vids)), "f_init": np.empty(n), "f_end": np.empty(n), "iters": np.empty(n, np.int32),
               "status": np.empty(n, np.int32), "n_feval": np.empty(n, np.int64), "n_geval": np.empty(n, np.int64)}
        self._ck(self._lib.rdisgpu_solve_cgd_csr(self._h, n, _p(ps.var_off), _p(ps.vids), _p(ps.fac_off), _p(ps.fids), _p(x0a),
                                                 maxiters, ftol, _p(out["x"]), _p(out["f_init"]), _p(out["f_end"]),
                                                 _p(out["iters"]), _p(out["status"]), _p(out["n_feval"]), _p(out["n_geval"])))
        return out

    def solve_lm(self, problems, x0=None, maxiters=25, ftol=3e-8, opts=None):
        """rdisgpu_solve_lm_csr (Levenberg-Marquardt, parity unpinned).  Returns dict(x, f_init, f_end, iters,
        stop, n_feval, n_jeval)."""
        ps = problems
        x0a = None if x0 is None else _arr(x0, np.float64)
        o = _arr(opts if opts is not None else [1e-3, 1e-15, 1e-15, ftol], np.float64)
        n = ps.n
        out = {"x": np.empty(len(ps.vids)), "f_init": np.empty(n), "f_end": np.empty(n), "iters": np.empty(n, np.int32),
               "stop": np.empty(n, np.int32), "n_feval": np.empty(n, np.int64), "n_jeval": np.empty(n, np.int64)}
        self._ck(self._lib.rdisgpu_solve_lm_csr(self._h, n, _p(ps.var_off), _p(ps.vids), _p(ps.fac_off), _p(ps.fids), _p(x0a),
                                                maxiters, _p(o), _p(out["x"]), _p(out["f_init"]), _p(out["f_end"]),
                                                _p(out["iters"]), _p(out["stop"]), _p(out["n_feval"]), _p(out["n_jeval"])))
        return out

    def solve_cgd_structs(self, problems, x0=None, maxiters=25, ftol=3e-8):
        """rdisgpu_solve_cgd (array of rdisgpu_problem / rdisgpu_result structs): same results."""
        ps = problems
        x0a = None if x0 is None else _arr(x0, np.float64)
        parr = problem_array(ps, x0a)
        xout = np.empty(len(ps.vids))
        rarr = (Result * ps.n)()
        for i in range(ps.n):
            rarr[i].x = xout.ctypes.data + 8 * int(ps.var_off[i])
        self._ck(self._lib.rdisgpu_solve_cgd(self._h, parr, ps.n, maxiters, ftol, rarr))
        return _unpack(rarr, ps.n, xout)

    def batch(self, problems):
        return Batch(self, problems)


def _unpack(rarr, n, xout):
    buf = np.frombuffer(rarr, dtype=np.dtype([("x", np.uint64), ("f_init", np.float64), ("f_end", np.float64),
                                              ("iters", np.int32), ("status", np.int32), ("n_feval", np.int64),
                                              ("n_geval", np.int64)]), count=n)
    return {"x": xout, "f_init": buf["f_init"].copy(), "f_end": buf["f_end"].copy(), "iters": buf["iters"].copy(),
            "status": buf["status"].copy(), "n_feval": buf["n_feval"].copy(), "n_geval": buf["n_geval"].copy()}


class Batch:
    """rdisgpu_batch: index lists resident in HBM, solve asynchronously, fetch later."""

    def __init__(self, ctx, problems):
        self.ctx = ctx
        self.ps = problems
        self._lib = ctx._lib
        h = C.c_void_p()
        ctx._ck(self._lib.rdisgpu_batch_create_csr(ctx._h, problems.n, _p(problems.var_off), _p(problems.vids),
                                                   _p(problems.fac_off), _p(problems.fids), C.byref(h)))
        self._h = h
        ctx._batches.add(self)
        self._rarr = (Result * problems.n)()
        self._xout = np.empty(len(problems.vids))
        for i in range(problems.n):
            self._rarr[i].x = self._xout.ctypes.data + 8 * int(problems.var_off[i])

    def solve(self, x0=None, maxiters=25, ftol=3e-8):
        """Enqueue the solve.  x0: host array (pinned for a truly async copy) or None."""
        if x0 is None:
            ptr = None
        elif isinstance(x0, np.ndarray):
            assert x0.dtype == np.float64 and x0.flags.c_contiguous
            self._x0_keep = x0
            ptr = C.c_void_p(x0.ctypes.data)
        else:  # raw host pointer (e.g. torch pinned tensor .data_ptr())
            ptr = C.c_void_p(int(x0))
        self.ctx._ck(self._lib.rdisgpu_batch_solve_cgd(self._h, ptr, maxiters, ftol))

    def fetch(self, want_x=True):
        s = C.c_double(0)
        if want_x:
            self.ctx._ck(self._lib.rdisgpu_batch_fetch(self._h, self._rarr, C.byref(s)))
            out = _unpack(self._rarr, self.ps.n, self._xout.copy())
        else:
            self.ctx._ck(self._lib.rdisgpu_batch_fetch(self._h, None, C.byref(s)))
            out = {}
        out["sum_f_end"] = s.value
        return out

    def objective_device(self, sum_dev_ptr):
        """*sum_dev += sum of f_end (device-side; asynchronous)."""
        self.ctx._ck(self._lib.rdisgpu_batch_objective_device(self._h, C.c_void_p(sum_dev_ptr)))

    def info(self):
        out = np.zeros(8, np.int32)
        self.ctx._ck(self._lib.rdisgpu_batch_info(self._h, _p(out)))
        d = dict(zip(("nprobs", "point_warps", "camera_blocks", "cluster_size", "camera_threads", "generic_problems",
                      "camera_nf_max", "last_launches"), out.tolist()))
        r = np.zeros(2, np.int32)
        self.ctx._ck(self._lib.rdisgpu_batch_resident_info(self._h, _p(r)))
        d["resident_problems"] = int(r[0]); d["resident_smem_bytes"] = int(r[1])
        return d

    @property
    def last_launches(self):
        return self._lib.rdisgpu_batch_last_launches(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rdisgpu_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
