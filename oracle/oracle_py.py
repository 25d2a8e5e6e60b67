"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; nothing under rdis_b200/ does.  See oracle/rdis_oracle.hpp for
what the oracle restates and how it is pinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_RESTATED = os.path.join(_HERE, "liboracle.so")
LIB_REFNRC = os.path.join(_HERE, "_ref", "liboracle_refnrc.so")
LIB_FMA = os.path.join(_HERE, "liboracle_fma.so")
LIB_DEVTRIG = os.path.join(_HERE, "liboracle_devtrig.so")
_TWINS = {"recip": "liboracle_recip.so", "treefold": "liboracle_treefold.so", "devtrig_treefold": "liboracle_devtrig_treefold.so"}

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the oracle (and, where /root/reference exists, the _ref variant)."""
    if force or not os.path.exists(LIB_RESTATED) or \
            os.path.getmtime(LIB_RESTATED) < max(os.path.getmtime(os.path.join(_HERE, f))
                                                  for f in ("oracle_capi.cpp", "rdis_oracle.hpp", "nr_minimize.hpp", "lm_oracle.hpp", "interval_oracle.hpp")):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    elif not os.path.exists(LIB_REFNRC) and os.path.exists("/root/reference/external/include/minimize_nrc.h"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _opt(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _load(path):
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.orc_create_nlpf.restype = vp
    lib.orc_create_nlpf.argtypes = [C.c_int64, _f64p, _f64p, C.c_int64, _i64p, _i32p, _f64p, _f64p, _u8p, _f64p]
    lib.orc_create_ba.restype = vp
    lib.orc_create_ba.argtypes = [C.c_int32, C.c_int32, _f64p, _f64p, C.c_int64, _i32p, _i32p, _f64p]
    lib.orc_load_bal.restype = vp
    lib.orc_load_bal.argtypes = [C.c_char_p, C.c_int64, C.c_int64]
    lib.orc_load_poly.restype = vp
    lib.orc_load_poly.argtypes = [C.c_char_p]
    lib.orc_make_sinusoid.restype = vp
    lib.orc_make_sinusoid.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int]
    lib.orc_destroy.argtypes = [vp]
    for name in ("orc_num_vars", "orc_num_factors", "orc_num_edges", "orc_ncams", "orc_npts"):
        getattr(lib, name).restype = C.c_int64
        getattr(lib, name).argtypes = [vp]
    lib.orc_kind.restype = C.c_int
    lib.orc_kind.argtypes = [vp]
    lib.orc_get_bounds.argtypes = [vp, _f64p, _f64p, _f64p, _f64p]
    lib.orc_get_xinit.restype = C.c_int
    lib.orc_get_xinit.argtypes = [vp, _f64p]
    lib.orc_export_nlpf.argtypes = [vp, _i64p, _i32p, _f64p, _f64p, _u8p, _f64p]
    lib.orc_export_ba.argtypes = [vp, _i32p, _i32p, _f64p]
    lib.orc_set_change_filter.argtypes = [C.c_int]
    lib.orc_set_x.argtypes = [vp, C.c_int64, vp, _f64p]
    lib.orc_get_x.argtypes = [vp, C.c_int64, vp, _f64p]
    lib.orc_unassign.argtypes = [vp, C.c_int64, vp]
    lib.orc_set_factor_const.argtypes = [vp, C.c_int64, _i64p, _f64p, _u8p]
    lib.orc_bounds.argtypes = [vp, _u8p, C.c_int64, vp, vp, vp, _f64p]
    lib.orc_eval.restype = C.c_double
    lib.orc_eval.argtypes = [vp, C.c_int64, vp, vp, C.c_int]
    lib.orc_grad.argtypes = [vp, C.c_int64, vp, C.c_int64, vp, _f64p]
    lib.orc_factor_grad.argtypes = [vp, C.c_int64, _f64p]
    lib.orc_solve_cgd.restype = C.c_double
    lib.orc_solve_cgd.argtypes = [vp, C.c_int64, _i32p, C.c_int64, _i64p, _f64p, C.c_int, C.c_double,
                                  C.POINTER(C.c_double), C.POINTER(C.c_int)]
    lib.orc_solve_cgd_batch.restype = C.c_double
    lib.orc_solve_cgd_batch.argtypes = [vp, C.c_int64, _i64p, _i32p, _i64p, _i64p, _f64p, C.c_int, C.c_double,
                                        _f64p, _f64p, _i32p, C.c_int, vp]
    lib.orc_solve_lm_batch.restype = C.c_double
    lib.orc_solve_lm_batch.argtypes = [vp, C.c_int64, _i64p, _i32p, _i64p, _i64p, _f64p, C.c_int, C.c_double,
                                       _f64p, _f64p, _i32p, _i32p]
    lib.orc_get_counters.argtypes = [vp, _i64p]
    lib.orc_reset_counters.argtypes = [vp]
    lib.orc_trace_enable.argtypes = [vp, C.c_int]
    lib.orc_trace_count.restype = C.c_int64
    lib.orc_trace_count.argtypes = [vp]
    lib.orc_trace_get.restype = C.c_int
    lib.orc_trace_get.argtypes = [vp, C.c_int64, _f64p, _f64p]
    lib.orc_variant.restype = C.c_char_p
    return lib


_LIBS = {}


def lib(variant="restated"):
    """variant: 'restated' (own-words NR driver), 'refnrc' (reference's minimize_nrc.h, oracle/_ref),
    'fma' (restated, compiled with FMA contraction: a perturbation twin), 'devtrig' (restated, sin/cos = the device's
    routines compiled for the host: what the strict device mode must match bit for bit), 'recip' / 'treefold' /
    'devtrig_treefold' (further perturbation twins, see oracle/Makefile)."""
    paths = {"restated": LIB_RESTATED, "refnrc": LIB_REFNRC, "fma": LIB_FMA, "devtrig": LIB_DEVTRIG}
    paths.update({k: os.path.join(_HERE, v) for k, v in _TWINS.items()})
    path = paths[variant]
    if path not in _LIBS:
        if not os.path.exists(path):
            if variant == "fma":
                subprocess.check_call(["make", "-C", _HERE, "fma"], stdout=subprocess.DEVNULL)
            elif variant in _TWINS:
                subprocess.check_call(["make", "-C", _HERE, "twins"], stdout=subprocess.DEVNULL)
            else:
                build(force=True)
        _LIBS[path] = _load(path)
    return _LIBS[path]


def have_refnrc():
    return os.path.exists(LIB_REFNRC)


class OracleFunction:
    """An oracle::OptimizableFunction plus its CGDSubspaceOptimizer."""

    def __init__(self, handle, variant):
        if not handle:
            raise RuntimeError("oracle: could not build / load the function")
        self._lib = lib(variant)
        self._h = C.c_void_p(handle)
        self.variant = variant
        self._spec = None

    def __del__(self):
        try:
            self._lib.orc_destroy(self._h)
        except Exception:
            pass

    # ---- constructors -------------------------------------------------------------
    @classmethod
    def nlpf(cls, lb, ub, rowptr, vid, expo, konst, sine, coeff, variant="restated"):
        L = lib(variant)
        lb = np.ascontiguousarray(lb, np.float64); ub = np.ascontiguousarray(ub, np.float64)
        h = L.orc_create_nlpf(len(lb), lb, ub, len(coeff), np.ascontiguousarray(rowptr, np.int64),
                              np.ascontiguousarray(vid, np.int32), np.ascontiguousarray(expo, np.float64),
                              np.ascontiguousarray(konst, np.float64), np.ascontiguousarray(sine, np.uint8),
                              np.ascontiguousarray(coeff, np.float64))
        return cls(h, variant)

    @classmethod
    def ba(cls, ncams, npts, lb, ub, cam, pt, obs, variant="restated"):
        L = lib(variant)
        h = L.orc_create_ba(ncams, npts, np.ascontiguousarray(lb, np.float64), np.ascontiguousarray(ub, np.float64),
                            len(cam), np.ascontiguousarray(cam, np.int32), np.ascontiguousarray(pt, np.int32),
                            np.ascontiguousarray(obs, np.float64).reshape(-1))
        return cls(h, variant)

    @classmethod
    def from_spec(cls, spec, variant="restated"):
        """spec: dict as produced by export() / rdis_b200.problems generators."""
        if spec["kind"] == "ba":
            return cls.ba(spec["ncams"], spec["npts"], spec["lb"], spec["ub"], spec["cam"], spec["pt"], spec["obs"], variant)
        return cls.nlpf(spec["lb"], spec["ub"], spec["rowptr"], spec["vid"], spec["expo"], spec["konst"],
                        spec["sine"], spec["coeff"], variant)

    @classmethod
    def load_bal(cls, path, ncams=0, npts=0, variant="restated"):
        return cls(lib(variant).orc_load_bal(path.encode(), ncams, npts), variant)

    @classmethod
    def load_poly(cls, path, variant="restated"):
        return cls(lib(variant).orc_load_poly(path.encode()), variant)

    @classmethod
    def sinusoid(cls, height, branches, arity, odd=False, variant="restated"):
        return cls(lib(variant).orc_make_sinusoid(height, branches, arity, int(odd)), variant)

    # ---- structure ----------------------------------------------------------------
    @property
    def V(self):
        return self._lib.orc_num_vars(self._h)

    @property
    def F(self):
        return self._lib.orc_num_factors(self._h)

    @property
    def E(self):
        return self._lib.orc_num_edges(self._h)

    def export(self):
        """Flat arrays describing the function (the C-ABI's input format)."""
        V, F, E = self.V, self.F, self.E
        lb = np.empty(V); ub = np.empty(V); slo = np.empty(V); shi = np.empty(V)
        self._lib.orc_get_bounds(self._h, lb, ub, slo, shi)
        spec = {"V": V, "F": F, "lb": lb, "ub": ub, "samp_lo": slo, "samp_hi": shi}
        x0 = np.empty(V)
        if self._lib.orc_get_xinit(self._h, x0):
            spec["x0"] = x0
        if self._lib.orc_kind(self._h) == 1:
            cam = np.empty(F, np.int32); pt = np.empty(F, np.int32); obs = np.empty(2 * F)
            self._lib.orc_export_ba(self._h, cam, pt, obs)
            spec.update(kind="ba", ncams=int(self._lib.orc_ncams(self._h)), npts=int(self._lib.orc_npts(self._h)),
                        cam=cam, pt=pt, obs=obs.reshape(F, 2))
        else:
            rowptr = np.empty(F + 1, np.int64); vid = np.empty(E, np.int32)
            expo = np.empty(E); konst = np.empty(E); sine = np.empty(E, np.uint8); coeff = np.empty(F)
            self._lib.orc_export_nlpf(self._h, rowptr, vid, expo, konst, sine, coeff)
            spec.update(kind="nlpf", rowptr=rowptr, vid=vid, expo=expo, konst=konst, sine=sine, coeff=coeff)
        return spec

    # ---- state --------------------------------------------------------------------
    def set_x(self, x, vid=None):
        x = np.ascontiguousarray(x, np.float64)
        v = _opt(vid, np.int32)
        self._lib.orc_set_x(self._h, len(x), _ptr(v), x)

    def get_x(self, vid=None):
        n = self.V if vid is None else len(vid)
        out = np.empty(n)
        v = _opt(vid, np.int32)
        self._lib.orc_get_x(self._h, n, _ptr(v), out)
        return out

    def unassign(self, vid=None):
        v = _opt(vid, np.int32)
        self._lib.orc_unassign(self._h, self.V if vid is None else len(v), _ptr(v))

    def set_factor_const(self, fid, val, on):
        self._lib.orc_set_factor_const(self._h, len(fid), np.ascontiguousarray(fid, np.int64),
                                       np.ascontiguousarray(val, np.float64), np.ascontiguousarray(on, np.uint8))

    def bounds(self, point, fid=None):
        """Factor::computeBounds per factor + the interval sum: (lower, upper, (sum_lower, sum_upper))."""
        pt = np.ascontiguousarray(point, np.uint8)
        f = _opt(fid, np.int64)
        n = self.F if fid is None else len(f)
        lo = np.empty(n); hi = np.empty(n); tot = np.zeros(2)
        self._lib.orc_bounds(self._h, pt, n, _ptr(f), _ptr(lo), _ptr(hi), tot)
        return lo, hi, (float(tot[0]), float(tot[1]))

    # ---- hot path -----------------------------------------------------------------
    def eval(self, fid=None, per_factor=False, use_cache=True):
        f = _opt(fid, np.int64)
        n = self.F if fid is None else len(f)
        pf = np.empty(n) if per_factor else None
        s = self._lib.orc_eval(self._h, n, _ptr(f), _ptr(pf), int(use_cache))
        return (s, pf) if per_factor else s

    def grad(self, fid=None, vid=None):
        f = _opt(fid, np.int64); v = _opt(vid, np.int32)
        nv = self.V if vid is None else len(v)
        g = np.empty(nv)
        self._lib.orc_grad(self._h, self.F if fid is None else len(f), _ptr(f), nv, _ptr(v), g)
        return g

    def factor_grad(self, fid, arity):
        out = np.empty(arity)
        self._lib.orc_factor_grad(self._h, fid, out)
        return out

    def solve_cgd(self, vid, fid, x0, maxiters=25, ftol=3e-8):
        """One CGDSubspaceOptimizer::optimize call -> (f_end, delta, x, iters)."""
        vid = np.ascontiguousarray(vid, np.int32); fid = np.ascontiguousarray(fid, np.int64)
        x = np.array(x0, dtype=np.float64, copy=True)
        d = C.c_double(0); it = C.c_int(0)
        fe = self._lib.orc_solve_cgd(self._h, len(vid), vid, len(fid), fid, x, maxiters, ftol, C.byref(d), C.byref(it))
        return fe, d.value, x, it.value

    def solve_cgd_batch(self, var_off, vids, fac_off, fids, x0, maxiters=25, ftol=3e-8, replicas=None):
        """Batch of solves; replicas = list of OracleFunction copies -> one thread each.
        Returns dict(x, f_end, f_init, iters, seconds)."""
        var_off = np.ascontiguousarray(var_off, np.int64); fac_off = np.ascontiguousarray(fac_off, np.int64)
        vids = np.ascontiguousarray(vids, np.int32); fids = np.ascontiguousarray(fids, np.int64)
        n = len(var_off) - 1
        x = np.array(x0, dtype=np.float64, copy=True)
        fe = np.empty(n); fi = np.empty(n); it = np.empty(n, np.int32)
        if replicas:
            arr = (C.c_void_p * len(replicas))(*[r._h for r in replicas])
            secs = self._lib.orc_solve_cgd_batch(self._h, n, var_off, vids, fac_off, fids, x, maxiters, ftol, fe, fi, it,
                                                 len(replicas), C.cast(arr, C.c_void_p))
        else:
            secs = self._lib.orc_solve_cgd_batch(self._h, n, var_off, vids, fac_off, fids, x, maxiters, ftol, fe, fi, it,
                                                 1, None)
        return {"x": x, "f_end": fe, "f_init": fi, "iters": it, "seconds": secs}

    def solve_lm_batch(self, var_off, vids, fac_off, fids, x0, maxiters=25, ftol=3e-8):
        """Batch of LMSubspaceOptimizer::optimize calls (PARITY UNPINNED: levmar restated, oracle/lm_oracle.hpp).
        Returns dict(x, f_end, f_init, iters, stop, seconds)."""
        var_off = np.ascontiguousarray(var_off, np.int64); fac_off = np.ascontiguousarray(fac_off, np.int64)
        vids = np.ascontiguousarray(vids, np.int32); fids = np.ascontiguousarray(fids, np.int64)
        n = len(var_off) - 1
        x = np.array(x0, dtype=np.float64, copy=True)
        fe = np.empty(n); fi = np.empty(n); it = np.empty(n, np.int32); st = np.empty(n, np.int32)
        secs = self._lib.orc_solve_lm_batch(self._h, n, var_off, vids, fac_off, fids, x, maxiters, ftol, fe, fi, it, st)
        return {"x": x, "f_end": fe, "f_init": fi, "iters": it, "stop": st, "seconds": secs}

    def trace(self, on=True):
        """Start (and clear) / stop recording every SubfunctionFD evaluation of the next solves."""
        self._lib.orc_trace_enable(self._h, int(on))

    def trace_records(self, nv):
        """[(is_df, x[nv], out[1 or nv])] of the recorded solve."""
        recs = []
        for i in range(self._lib.orc_trace_count(self._h)):
            x = np.empty(nv); out = np.empty(nv)
            is_df = self._lib.orc_trace_get(self._h, i, x, out)
            recs.append((is_df, x, out if is_df else out[:1].copy()))
        return recs

    def counters(self):
        out = np.zeros(5, np.int64)
        self._lib.orc_get_counters(self._h, out)
        return dict(zip(("factor_eval_calls", "factor_recomputes", "factor_grad_calls", "f_evals", "df_evals"), out.tolist()))

    def reset_counters(self):
        self._lib.orc_reset_counters(self._h)


def set_change_filter(on, variant="restated"):
    lib(variant).orc_set_change_filter(int(on))
