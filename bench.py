#!/usr/bin/env python
"""bench.py — subspace-solves/sec on ladybug-49-7776 (BASELINE.json's metric and graph).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

One *step* = one alternating wave of the recursive decomposer's leaf work on the reference's own data file
data/ladybug-problem-49-7776-pre.txt (49 cameras, 7776 points, 31843 observations; committed as the fixture
tests/golden/ladybug_49_7776.npz), from the file's state (`--randinit 0`): reset the state to x0, solve the 7776
point components given the cameras (one sibling batch), then the 49 camera components given the points (second
sibling batch) — 7825 CGDSubspaceOptimizer::optimize calls (SSmaxit 25, ftol 3e-8, the optBA defaults) — and
accumulate the global objective.

  value   solves/s, device-resident: index lists and x0 already in HBM, results stay in HBM
  e2e     N=1: the same step through the C++ plugin a reference maintainer binds — rdis::CudaSubspaceOptimizer::
          optimizeBatch over rdis::Variable / rdis::Factor host objects (tests/native/host_driver benchwaves):
          Variable::assign of the start state, index lists + start values up, results down, Variable write-back.
          N>1: the sharded step through the C-ABI (rdisgpu_solve_cgd_csr) with host buffers and a host all-gather.
  roofline        the dominant kernel of the step, algorithmic bytes (every objective evaluation =
                  32 B/factor + 8 B/variable touched, +8 B/variable with gradients) over its CUDA-event time; the
                  solves are L2-resident and latency / FP64 bound
  roofline_sweep  the residual sweep on the cfg4 graph (1,048,575 vars / 4,194,292 factors, 268 MB per sweep —
                  larger than L2): the HBM-roofline evidence for the sweep kernel
  cpu_baseline    the CPU oracle (reference-semantics port, 1 core) on a bounded sample of the wave

Multi-GPU = STRONG scaling of the ONE sibling set (north_star / BASELINE config 5): the components of each wave are
dealt to the ranks (LPT by factor count), every rank keeps the whole graph resident, the point wave's results are
all-gathered (NCCL) into every replica before the camera wave, the objective is all-reduced.  `replicas_weak` keeps
last round's weak-scaling figure (one whole set per rank) as a secondary key.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAXITERS, FTOL = 25, 3e-8
WORKLOAD = ("optBA ladybug-49-7776 (reference data file: 49 cams, 7776 pts, 31843 obs), file state: one alternating wave = "
            "7776 point-component + 49 camera-component CGDSubspaceOptimizer solves per step")
DATA = "reference data/ladybug-problem-49-7776-pre.txt (fixture tests/golden/ladybug_49_7776.npz); synthetic graphs for cfg2 / cfg4 legs"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-sweep", action="store_true", help="skip the secondary legs (cfg2 / cfg3-full / cfg4 / sweeps / LM)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and parity legs")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU budget of the cpu_baseline sample")
    return ap.parse_args()


def config(world):
    """Identical in both arms at a given N (the driver compares them)."""
    return {"workload": WORKLOAD, "ssmaxit": MAXITERS, "ssftol": FTOL, "n_gpus": world,
            "parallelism": "GPU arm: component shard x%d of ONE sibling set (LPT by factor count; graph replicated on every GPU, point "
                           "results all-gathered before the camera wave, objective all-reduced), strong scaling.  CPU arm: all host "
                           "threads, one function replica per thread, same total work at every N" % world,
            "l2": "GPU arm: flushed between steps (256 MiB memset + read pass, untimed); working set 1.4 MB.  CPU arm: n/a"}


def load_problems_module():
    """rdis_b200/problems.py by path: numpy only, so the reference arm never maps librdis_b200.so."""
    import importlib.util
    sp = importlib.util.spec_from_file_location("rdis_problems", os.path.join(ROOT, "rdis_b200", "problems.py"))
    mod = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs: ONE nvidia-smi process looping every
    50 ms (spawning one per sample costs more than the timed region lasts)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.25)  # first samples are on their way before the timed region starts
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            for line in out.strip().splitlines():
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 7:
                    rows.append(parts)
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(rows)}


def algorithmic_bytes(ps, res):
    """one objective evaluation over a problem = 32 B per factor + 8 B per distinct
    variable it touches; evaluations that also produce derivatives add 8 B per variable."""
    nf = np.diff(ps.fac_off)
    nv = np.diff(ps.var_off)
    if ps.n and nv[0] == 3:      # point components: 3 own variables + 9 per observing camera
        touched = 3 + 9 * nf
    else:                        # camera components: 9 own variables + 3 per observed point
        touched = 9 + 3 * nf
    n_f = res["n_feval"].astype(np.float64)
    n_g = res["n_geval"].astype(np.float64)
    return float(np.sum(n_f * (32.0 * nf + 8.0 * touched) + n_g * 8.0 * touched))


def measured_traffic(kernel, key="dram_bytes"):
    """DRAM bytes per launch (or another recorded metric) of `kernel` from the committed ncu captures
    (profiles/r02_traffic.json, falling back to round 1's), or None."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                v = json.load(f).get(kernel, {}).get(key)
            if v is not None:
                return v
        except Exception:
            pass
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def write_bal(spec, path):
    """The fixture back in the reference's BAL text format (17 significant digits: the doubles round-trip), for the
    C++ loader of the plugin (rdis::BundleAdjustmentFunction::load)."""
    nc, npt, F = int(spec["ncams"]), int(spec["npts"]), int(spec["F"])
    with open(path, "w") as f:
        f.write("%d %d %d\n" % (nc, npt, F))
        obs = spec["obs"].reshape(F, 2)
        f.write("".join("%d %d %.17g %.17g\n" % (c, p, o[0], o[1]) for c, p, o in zip(spec["cam"], spec["pt"], obs)))
        f.write("".join("%.17g\n" % v for v in spec["x0"]))


def lpt_shard(ps, world):
    """Deterministic LPT packing of a wave's components onto the ranks by factor count (every evaluation costs one
    pass over the factors; the evaluation counts are not known before the solve)."""
    cost = np.diff(ps.fac_off).astype(np.int64)
    order = np.argsort(-cost, kind="stable")
    owner = np.empty(ps.n, dtype=np.int64)
    load = np.zeros(world, dtype=np.int64)
    if ps.n > 64 * world:
        owner[order] = np.arange(ps.n) % world          # thousands of small components: dealing the sorted list is LPT-tight
    else:
        for i in order:
            r = int(np.argmin(load))
            owner[i] = r
            load[r] += cost[i]
    return [np.nonzero(owner == r)[0] for r in range(world)]


class DevView:
    """A raw device pointer as a torch tensor (zero copy) through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from rdis_b200 import Context, problems as P

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")   # the host-buffer exchange of the e2e leg

    # ---- workload: the ONE sibling set of the real graph, its components dealt to the ranks ----
    spec = P.load_golden_ba()
    x0 = spec["x0"]
    V = spec["V"]
    pts, cams = P.ba_point_problems(spec), P.ba_camera_problems(spec)
    n_solves = pts.n + cams.n
    own_p, own_c = lpt_shard(pts, world), lpt_shard(cams, world)
    my_pts = pts.subset(own_p[rank]) if world > 1 else pts
    my_cams = cams.subset(own_c[rank]) if world > 1 else cams
    stream = torch.cuda.current_stream()
    ctx = Context.from_spec(spec, device=local_rank, stream=stream.cuda_stream)
    x0_dev = torch.from_numpy(x0).to(dev)
    obj_dev = torch.zeros(1, dtype=torch.float64, device=dev)
    b_pts, b_cams = ctx.batch(my_pts), ctx.batch(my_cams)
    flush = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    # exchange of the point wave: every rank's solved point values -> every replica (all-gather of equal-sized slices;
    # the pad repeats a rank's first variable, so the scatter writes the same value twice)
    state_ptr, _ = ctx.device_state()
    state = torch.as_tensor(DevView(state_ptr, (V, 2)), device=dev)
    if world > 1:
        nmax = max(len(pts.subset(o).vids) for o in own_p)
        pad = lambda v: np.concatenate([v, np.full(nmax - len(v), v[0], v.dtype)])
        vid_all = torch.from_numpy(np.concatenate([pad(pts.subset(o).vids) for o in own_p]).astype(np.int32)).to(dev)
        my_vid = torch.from_numpy(pad(my_pts.vids).astype(np.int64)).to(dev)
        recv = torch.empty(world * nmax, dtype=torch.float64, device=dev)

    def step_resident(timers=None):
        obj_dev.zero_()
        ctx.set_x_device(x0_dev.data_ptr(), V)
        b_pts.solve(None, MAXITERS, FTOL)
        b_pts.objective_device(obj_dev.data_ptr())
        if timers is not None:
            timers[1].record(stream)
        if world > 1:
            send = state[:, 0].index_select(0, my_vid)
            dist.all_gather_into_tensor(recv, send)
            ctx.set_x_device(recv.data_ptr(), world * nmax, vid_all.data_ptr())
        if timers is not None:
            timers[2].record(stream)
        b_cams.solve(None, MAXITERS, FTOL)
        b_cams.objective_device(obj_dev.data_ptr())
        if world > 1:
            dist.all_reduce(obj_dev)

    launches_per_step = None
    for _ in range(max(args.warmup, 3)):
        l0 = ctx.launch_count
        step_resident()
        launches_per_step = ctx.launch_count - l0
    torch.cuda.synchronize()

    # ---- timed region: K steps, CUDA events on the launching stream, L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                     # untimed: evicts the previous step's working set from L2 ...
        flush.sum()                       # ... and a read pass leaves clean lines (no write-backs inside the timing)
        ev[k][0].record(stream)
        step_resident(ev[k])
        ev[k][3].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    ms_steps = np.array([e[0].elapsed_time(e[3]) for e in ev])
    ms_pts = np.array([e[0].elapsed_time(e[1]) for e in ev])
    ms_xchg = np.array([e[1].elapsed_time(e[2]) for e in ev])
    ms_cams = np.array([e[2].elapsed_time(e[3]) for e in ev])
    total_ms = float(ms_steps.sum())
    per_rank = None
    if world > 1:
        t = torch.tensor([total_ms, float(ms_pts.mean()), float(ms_xchg.mean()), float(ms_cams.mean())], dtype=torch.float64, device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [[float(v) for v in a.tolist()] for a in allt]
        total_ms = max(a[0] for a in per_rank)
    sum_all = float(obj_dev.item())   # sum of f_end over the 7825 solves of the last step (all ranks): point wave + camera wave
    value = n_solves * args.steps / (total_ms * 1e-3)

    # ---- results of the last step (for the roofline's evaluation counts and the residual-eval rate) ----
    r_pts, r_cams = b_pts.fetch(), b_cams.fetch()
    evals = float(np.sum(r_pts["n_feval"] * np.diff(my_pts.fac_off)) + np.sum(r_cams["n_feval"] * np.diff(my_cams.fac_off)))
    if world > 1:
        t = torch.tensor([evals], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        evals = float(t.item())
    resid_evals_per_s = evals * args.steps / (total_ms * 1e-3)
    # the objective after the step = the camera wave's final values (every factor belongs to exactly one camera component)
    t = torch.tensor([float(r_cams["f_end"].sum()), float(r_pts["f_end"].sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    objective, pts_sum = float(t[0].item()), float(t[1].item())

    # ---- the same step without the launch-order history (what the FIRST visit of a sibling set costs) ----
    first_visit = None
    if world == 1:
        ctx.set_option("adaptive_order", 0)
        b1p, b1c = ctx.batch(my_pts), ctx.batch(my_cams)
        ts = []
        for _ in range(6):
            flush.zero_()
            flush.sum()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            ctx.set_x_device(x0_dev.data_ptr(), V)
            b1p.solve(None, MAXITERS, FTOL)
            b1c.solve(None, MAXITERS, FTOL)
            e.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        first_visit = {"ms_per_step": float(np.mean(ts[1:])), "what": "launch order as built (size classes, problem order): no evaluation "
                       "counts of a previous visit to sort by (rdisgpu_set_option adaptive_order = 0); results identical"}
        b1p.close()
        b1c.close()
        ctx.set_option("adaptive_order", 1)

    # ---- e2e ----
    e2e = e2e_leg(args, spec, pts, cams, own_p, own_c, ctx, rank, world, local_rank, host_group, dist, torch, dev)

    out = None
    if rank == 0:
        peak, peak_src = peaks()
        dom_is_pts = ms_pts.mean() >= ms_cams.mean()
        dom_ps, dom_res, dom_ms = (my_pts, r_pts, ms_pts.mean()) if dom_is_pts else (my_cams, r_cams, ms_cams.mean())
        dom_name = "solve_ba_points_kernel" if dom_is_pts else "solve_ba_cameras_kernel"
        abytes = algorithmic_bytes(dom_ps, dom_res)
        achieved = abytes / (dom_ms * 1e-3) / 1e9
        cfg = config(world)
        out = {
            "metric": "subspace-solves/sec", "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": DATA, "config": cfg,
            "residual_evals_per_sec": resid_evals_per_s,
            "objective_after_step": objective, "point_wave_sum_f_end": pts_sum, "sum_f_end_all_solves_device": sum_all,
            "mapping": {"points": b_pts.info(), "cameras": b_cams.info()},
            "kernel_ms": {"solve_ba_points_kernel": float(ms_pts.mean()), "solve_ba_cameras_kernel": float(ms_cams.mean()),
                          "exchange_allgather_scatter": float(ms_xchg.mean())},
            "per_rank_ms": None if per_rank is None else {"columns": ["total over the timed steps", "points wave", "exchange", "cameras wave"], "rows": per_rank},
            "limiting": "the longest line-search chain of one camera component (cameras wave) — per-rank work shrinks with N, the chain does not",
            "wall_s_timed_region": t_wall,
            "first_visit": first_visit,
            "launch_order": "block-kernel launch order re-sorted on the device after every solve by the evaluation counts just observed "
                            "(longest chains first): the timed steps are revisits of the same sibling set, as in the tree search's "
                            "alternating minimisation (src/RDISOptimizer.cpp:1148-1181); first_visit = without that history",
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "e2e": e2e,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(dom_name), "kernel": dom_name,
                         "algorithmic_bytes_per_launch": abytes, "launch_ms": float(dom_ms), "peak_source": peak_src,
                         "fp64_pipe_active_pct_ncu": measured_traffic(dom_name, "fp64_pipe_active_pct"),
                         "note": "working set 1.4 MB: L2-resident, a serial chain of <=1000 dependent evaluations per problem: "
                                 "latency bound, HBM is not the limiter here; see roofline_sweep for the HBM-bound kernel"},
        }
    # ---- secondary legs ----
    if not args.no_sweep:
        rw = replicas_weak(spec, pts, cams, local_rank, stream, dev, world, dist, torch)
        c4 = cfg4_sibling_wave(local_rank, stream, rank, world, dist if world > 1 else None, dev)
        if rank == 0:
            out["replicas_weak"] = rw
            out["cfg4_sibling_wave"] = c4
            out["cfg4_sibling_wave_ms"] = c4["ms"]
    if rank == 0 and not args.no_sweep:
        out["lm_wave"] = lm_wave(ctx, spec, pts, cams, x0, torch)
        out["ba_residual_sweep"] = ba_sweep_rates(spec, local_rank, stream, dev)
        out["batched_samples"] = batched_samples(spec, local_rank, stream, dev, torch)
        out["cfg2_full_solve"] = cfg2_full_solve(local_rank, stream, with_cpu=not args.no_cpu)
        out["cfg3_full"] = cfg3_full(spec, local_rank, stream, with_cpu=not args.no_cpu)
        out["roofline_sweep"] = sweep_roofline(local_rank, stream)
    if rank == 0 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(spec, pts, cams, x0, args.cpu_seconds)
        out["parity"] = parity_leg(spec, pts, cams, local_rank, stream)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def e2e_leg(args, spec, pts, cams, own_p, own_c, ctx, rank, world, local_rank, host_group, dist, torch, dev):
    x0 = spec["x0"]
    V = spec["V"]
    n_solves = pts.n + cams.n
    e2e_steps = max(3, min(args.steps, 10))
    # bytes per step.  up: the start state (8 B/variable), per batch one index blob (24 B ProblemDesc + 3 x 4 B order lists +
    # 12 B warp task per problem, 4 B per variable id and factor id) + start values; down: 32 B result record per problem +
    # final values.  The plugin uploads (4 B id + 8 B value) for the variables the host changed instead of the dense state.
    h2d = 8 * V + sum(4 * len(p.vids) + 4 * len(p.fids) + 8 * len(p.vids) + 48 * p.n for p in (pts, cams))
    d2h = sum(8 * len(p.vids) + 32 * p.n for p in (pts, cams))
    if world == 1:
        # through the plugin: (4 B id + 8 B value) per variable the host assigned + the start values of every problem; the
        # index lists of a revisited sibling set stay resident on the device (the adapter's batch cache; built in the warm-up)
        res = {"value": None, "unit": "solves/s", "h2d_bytes_per_step": int(12 * V + sum(8 * len(p.vids) for p in (pts, cams))),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps}
        drv = os.path.join(ROOT, "tests", "native", "host_driver")
        with tempfile.TemporaryDirectory() as td:
            bal = os.path.join(td, "ladybug.txt")
            write_bal(spec, bal)
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local_rank)))
            p = subprocess.run([drv, "benchwaves", bal, str(e2e_steps), "3"], capture_output=True, text=True, env=env)
        if p.returncode != 0:
            raise RuntimeError("host_driver benchwaves failed: " + p.stderr[-400:])
        d = json.loads(p.stdout.strip().splitlines()[-1])
        res.update({"value": d["solves_per_s"], "ms_per_step": d["ms_per_step"], "objective": d["objective_after_step"],
                    "through": "rdis::CudaSubspaceOptimizer::optimizeBatch over host Variable/Factor objects (librdis_host.so, "
                               "tests/native/host_driver benchwaves): Variable::assign of x0, 2 sibling batches, Variable write-back; wall clock",
                    "dispatch_ms_once": d["dispatch_ms"], "plugin_ms_per_step": d.get("plugin_ms_per_step")})
        # the same step through the bare C-ABI from ctypes (what round 1 reported as e2e)
        x0_pts, x0_cams = x0[pts.vids].copy(), x0[cams.vids].copy()

        def step_capi():
            ctx.set_x(x0)
            ra = ctx.solve_cgd(pts, x0_pts, MAXITERS, FTOL)
            xc = ctx.get_x(cams.vids)
            rb = ctx.solve_cgd(cams, xc, MAXITERS, FTOL)
            return float(rb["f_end"].sum())
        for _ in range(2):
            step_capi()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            obj = step_capi()
        t = time.perf_counter() - t0
        res["c_abi_ctypes"] = {"value": n_solves * e2e_steps / t, "ms_per_step": t / e2e_steps * 1e3, "objective": obj}
        return res
    # N > 1: every rank solves its shard through rdisgpu_solve_cgd_csr with host buffers; the point results travel
    # host-to-host (gloo all-gather) and are uploaded into every replica before the camera wave
    my_pts, my_cams = pts.subset(own_p[rank]), cams.subset(own_c[rank])
    x0_pts = x0[my_pts.vids].copy()
    nmax = max(len(pts.subset(o).vids) for o in own_p)
    vid_all = np.concatenate([np.concatenate([pts.subset(o).vids, np.full(nmax - len(pts.subset(o).vids), pts.subset(o).vids[0], np.int32)]) for o in own_p])
    send = torch.zeros(nmax, dtype=torch.float64)
    recv = torch.zeros(world * nmax, dtype=torch.float64)

    def step():
        ctx.set_x(x0)
        ra = ctx.solve_cgd(my_pts, x0_pts, MAXITERS, FTOL)
        send[:len(ra["x"])] = torch.from_numpy(ra["x"])
        send[len(ra["x"]):] = float(ra["x"][0])
        dist.all_gather_into_tensor(recv, send, group=host_group)
        ctx.set_x(recv.numpy(), vid_all)
        rb = ctx.solve_cgd(my_cams, None, MAXITERS, FTOL)
        part = torch.tensor([float(rb["f_end"].sum())], dtype=torch.float64)
        dist.all_reduce(part, group=host_group)
        return float(part.item())
    for _ in range(2):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        obj = step()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_e2e = float(t.item())
    return {"value": n_solves * e2e_steps / t_e2e, "unit": "solves/s", "h2d_bytes_per_step": int(h2d + 8 * V * (world - 1) + 12 * world * nmax),
            "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": t_e2e / e2e_steps * 1e3, "objective": obj,
            "through": "rdisgpu_solve_cgd_csr with host buffers on every rank (ctypes), point results all-gathered host-to-host (gloo), "
                       "objective all-reduced; wall clock, max over ranks"}


def replicas_weak(spec, pts, cams, device, stream, dev, world, dist, torch):
    """Last round's multi-GPU figure, kept as a secondary key: every rank solves one WHOLE copy of the sibling set
    (weak scaling; objective all-reduced)."""
    from rdis_b200 import Context
    ctx = Context.from_spec(spec, device=device, stream=stream.cuda_stream)
    x0_dev = torch.from_numpy(spec["x0"]).to(dev)
    obj = torch.zeros(1, dtype=torch.float64, device=dev)
    bp, bc = ctx.batch(pts), ctx.batch(cams)
    steps = 10
    ms = []
    for it in range(steps + 3):
        obj.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        a.record(stream)
        ctx.set_x_device(x0_dev.data_ptr(), spec["V"])
        bp.solve(None, MAXITERS, FTOL)
        bp.objective_device(obj.data_ptr())
        bc.solve(None, MAXITERS, FTOL)
        bc.objective_device(obj.data_ptr())
        if world > 1:
            dist.all_reduce(obj)
        e.record(stream)
        torch.cuda.synchronize()
        if it >= 3:
            ms.append(a.elapsed_time(e))
    t = torch.tensor([float(np.sum(ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = float(t.item())
    return {"what": "one whole copy of the ladybug sibling set per rank (weak scaling), objective all-reduced", "n_gpus": world,
            "value": world * (pts.n + cams.n) * steps / (tot * 1e-3), "unit": "solves/s", "ms_per_step": tot / steps, "scaling": "weak"}


def top_level_block(spec, P):
    """A top-level block of round(0.2 V) variables (src/RDISOptimizer.cpp:1755-1764): 10 camera blocks + 1555 point blocks
    = 4755 variables, with every factor all of whose variables are then assigned."""
    ncams = spec["ncams"]
    cam_sel, pt_sel = np.arange(10), np.arange(1555)
    vids = np.sort(np.concatenate([(9 * cam_sel[:, None] + np.arange(9)).ravel(), (9 * ncams + 3 * pt_sel[:, None] + np.arange(3)).ravel()])).astype(np.int32)
    fids = np.nonzero(np.isin(spec["cam"], cam_sel) | np.isin(spec["pt"], pt_sel))[0].astype(np.int64)
    return P.ProblemSet([0, len(vids)], vids, [0, len(fids)], fids)


def lm_wave(ctx, spec, pts, cams, x0, torch):
    """The same wave with the Levenberg-Marquardt subspace solver (BASELINE config 3: per-component LM); host buffers
    through rdisgpu_solve_lm_csr."""
    x0p, x0c = x0[pts.vids].copy(), x0[cams.vids].copy()

    def step(maxit):
        ctx.set_x(x0)
        ra = ctx.solve_lm(pts, x0p, maxit, FTOL)
        rb = ctx.solve_lm(cams, x0c, maxit, FTOL)
        return float(ra["f_end"].sum() + rb["f_end"].sum()), ra, rb
    res = {}
    for label, maxit in (("ssmaxit_25", MAXITERS), ("itmax_200", 200)):
        for _ in range(2):
            step(maxit)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            obj, a, b = step(maxit)
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / 3
        res[label] = {"value": (pts.n + cams.n) / t, "unit": "solves/s (host buffers, rdisgpu_solve_lm_csr)", "ms_per_step": t * 1e3,
                      "objective": obj, "iters_mean": float(np.concatenate([a["iters"], b["iters"]]).mean()),
                      "stop_histogram": np.bincount(np.concatenate([a["stop"], b["stop"]]), minlength=8).tolist()}
    # LM on the component RDIS poses FIRST on ladybug (the 4755-variable top-level block): lm_dense.cuh — block-sparse
    # Jacobian rows, dense 4755 x 4755 normal equations, blocked Cholesky with the trailing update on FP64 tensor cores
    from rdis_b200 import problems as P
    blk = top_level_block(spec, P)
    xs = x0[blk.vids].copy()
    for label, maxit in (("top_level_block_4755_vars_ssmaxit_25", MAXITERS), ("top_level_block_4755_vars_itmax_100", 100)):
        ctx.set_x(x0)
        ctx.solve_lm(blk, xs, 2, FTOL)
        ctx.set_x(x0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = ctx.solve_lm(blk, xs, maxit, FTOL)
        t = time.perf_counter() - t0
        res[label] = {"nv": int(len(blk.vids)), "nf": int(len(blk.fids)), "host_call_ms": t * 1e3, "f_init": float(r["f_init"][0]), "f_end": float(r["f_end"][0]),
                      "iters": int(r["iters"][0]), "stop": int(r["stop"][0]), "func_evals": int(r["n_feval"][0]), "jac_evals": int(r["n_jeval"][0]),
                      "ms_per_iteration": t * 1e3 / max(int(r["iters"][0]), 1),
                      "kernels": "lm_rows / lm_assemble / lm_potrf / lm_trsm / lm_syrk (DMMA) / lm_trsv; the reference hands levmar a dense "
                                 "11950 x 4755 Jacobian (454 MB) and an O(m^3) LU per damping trial on one core"}
    res["stop_codes"] = "levmar: 1 small gradient, 2 small step, 3 itmax, 4 singular, 5 no further reduction, 6 small ||e||, 7 non-finite"
    res["parity"] = "unpinned upstream (levmar not vendored); tested against oracle/lm_oracle.hpp"
    return res


def batched_samples(spec, device, stream, dev, torch):
    """Throughput regime: 8 independent copies of the problem (optBA's --nsamples restarts) in ONE batch."""
    from rdis_b200 import Context, problems as P
    K = 8
    spec_k = P.ba_replicate(spec, K)
    pts_k, cams_k = P.ba_point_problems(spec_k), P.ba_camera_problems(spec_k)
    ctx_k = Context.from_spec(spec_k, device=device, stream=stream.cuda_stream)
    xk = torch.from_numpy(spec_k["x0"]).to(dev)
    bpk, bck = ctx_k.batch(pts_k), ctx_k.batch(cams_k)
    tk = []
    for it in range(5):
        ctx_k.set_x_device(xk.data_ptr(), spec_k["V"])
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        bpk.solve(None, MAXITERS, FTOL)
        bck.solve(None, MAXITERS, FTOL)
        e.record(stream)
        torch.cuda.synchronize()
        tk.append(a.elapsed_time(e))
    ms_k = float(np.median(tk[2:]))
    return {"what": "%d independent copies of the ladybug problem in one graph: %d point + %d camera components per wave (NOT the "
                    "headline workload: shows the throughput regime of the same kernels)" % (K, pts_k.n, cams_k.n),
            "ms_per_wave": ms_k, "solves_per_sec": (pts_k.n + cams_k.n) / (ms_k * 1e-3), "camera_mapping": bck.info()}


def ba_sweep_rates(spec, device, stream, dev):
    """residual-evals/sec of the all-factor BA sweep (evalFactors over the whole graph) and of the residual +
    Jacobian-rows sweep (the LM path's 128 F + 8 V bytes): at ladybug size both are launch-bound (1.2 / 4.3 MB); 100
    copies of the point cloud (> L2) show the streaming regime."""
    import torch
    from rdis_b200 import Context, problems as P
    peak, _ = peaks()
    res = {}
    for name, K in (("ladybug", 1), ("x100_points_same_49_cameras", 100), ("x100_independent_copies_4900_cameras", -100)):
        sp = spec if K == 1 else (P.ba_replicate_points(spec, K) if K > 0 else P.ba_replicate(spec, -K))
        ctx = Context.from_spec(sp, device=device, stream=stream.cuda_stream)
        ctx.set_x(sp["x0"])
        pf = torch.empty(sp["F"], dtype=torch.float64, device=dev)
        tot = torch.zeros(1, dtype=torch.float64, device=dev)
        for _ in range(3):
            ctx.eval_device(tot.data_ptr(), pf.data_ptr())
        reps = 20
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            ctx.eval_device(tot.data_ptr(), pf.data_ptr())
        e.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(e) / reps
        abytes = 32.0 * sp["F"] + 8.0 * sp["V"]
        res[name] = {"factors": int(sp["F"]), "variables": int(sp["V"]), "ms_per_sweep": ms, "residual_evals_per_sec": sp["F"] / (ms * 1e-3),
                     "algorithmic_bytes": abytes, "achieved_GBps": abytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": abytes / (ms * 1e-3) / 1e9 / peak,
                     "sum": float(tot.item())}
        if hasattr(ctx, "factor_rows_device"):
            rows = torch.empty(sp["F"] * 12, dtype=torch.float64, device=dev)
            for _ in range(3):
                ctx.factor_rows_device(pf.data_ptr(), rows.data_ptr())
            torch.cuda.synchronize()
            a.record(stream)
            for _ in range(reps):
                ctx.factor_rows_device(pf.data_ptr(), rows.data_ptr())
            e.record(stream)
            torch.cuda.synchronize()
            ms = a.elapsed_time(e) / reps
            rb = 128.0 * sp["F"] + 8.0 * sp["V"]
            res[name]["jacobian_rows"] = {"ms_per_sweep": ms, "algorithmic_bytes": rb, "achieved_GBps": rb / (ms * 1e-3) / 1e9,
                                          "frac_of_hbm_peak": rb / (ms * 1e-3) / 1e9 / peak, "residual_evals_per_sec": sp["F"] / (ms * 1e-3)}
        del ctx
    res["kernels"] = "ba_camera_table_kernel + ba_sweep_kernel (back-to-back launches, per-factor values written)"
    return res


def cfg2_full_solve(device, stream, with_cpu=True):
    """BASELINE config 2 (optSinusoid d=1000): the chain h=999 k=1 arity 3 (V=1000, F=2999) and the
    default-shaped tree h=6 k=3 arity 3 (V=1093, F=3278), every variable and every factor in ONE subspace problem, SSmaxit 25,
    ftol 3e-8, seeded start.  One CTA solves it resident in shared memory (nlpf_resident.cuh); the CPU oracle solves
    the same problem from the same start on one core."""
    import torch
    from rdis_b200 import Context, problems as P
    res = {}
    for name, (h, k, ar) in (("chain_h999_k1_arity3", (999, 1, 3)), ("tree_h6_k3_arity3", (6, 3, 3))):
        spec = P.sinusoid(h, k, ar)
        x0 = P.random_start(spec, 834725927 % 1000)
        ps = P.full_problem(spec)
        ctx = Context.from_spec(spec, device=device, stream=stream.cuda_stream)
        b = ctx.batch(ps)
        x0_dev = torch.from_numpy(x0).to("cuda:%d" % device)
        ts = []
        for it in range(4):
            ctx.set_x_device(x0_dev.data_ptr(), spec["V"])
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            b.solve(None, MAXITERS, FTOL)
            e.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        r = b.fetch()
        row = {"V": int(spec["V"]), "F": int(spec["F"]), "gpu_ms": float(min(ts[1:])), "f_init": float(r["f_init"][0]),
               "f_end": float(r["f_end"][0]), "iters": int(r["iters"][0]), "evaluations": int(r["n_feval"][0]), "mapping": b.info(),
               "status": int(r["status"][0])}
        if with_cpu:
            from oracle import oracle_py as O
            orc = O.OracleFunction.from_spec(spec)
            orc.set_x(x0)
            o = orc.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0, MAXITERS, FTOL)
            row.update({"cpu_oracle_ms": o["seconds"] * 1e3, "f_end_cpu": float(o["f_end"][0]),
                        "rel_diff": float(abs(r["f_end"][0] - o["f_end"][0]) / abs(o["f_end"][0])),
                        "speedup_vs_1_core": o["seconds"] * 1e3 / float(min(ts[1:]))})
            sc = Context.from_spec(spec, device=device, stream=stream.cuda_stream)
            sc.set_option("strict", 1)
            sc.set_x(x0)
            od = O.OracleFunction.from_spec(spec, "devtrig")
            od.set_x(x0)
            rs = sc.solve_cgd(ps, x0, MAXITERS, FTOL)
            odr = od.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0, MAXITERS, FTOL)
            row["strict_bit_identical_to_devtrig_oracle"] = bool(rs["f_end"][0] == odr["f_end"][0] and np.array_equal(rs["x"], odr["x"]))
        res[name] = row
        b.close()
        del ctx
    return res


def cfg3_full(spec, device, stream, with_cpu=True):
    """BASELINE config 3 shape (i): the solves RDIS poses FIRST on ladybug (src/RDISOptimizer.cpp:1049-1067,1755-1764) — a
    top-level block of round(0.2 V) variables (10 camera blocks + 1555 point blocks = 4755 variables, every factor whose
    variables are assigned) and the whole graph as one problem (V = 23769, F = 31843), through the cooperative-grid
    kernel.  Timed at SSmaxit 25; parity against the CPU oracle at SSmaxit 1 (one line search: the CPU needs ~1 s per
    evaluation of the whole graph), production and strict mode."""
    import torch
    from rdis_b200 import Context, problems as P
    x0 = spec["x0"]
    block = top_level_block(spec, P)
    res = {}
    for name, ps in (("top_level_block_4755_vars", block), ("whole_graph_23769_vars", P.full_problem(spec))):
        ctx = Context.from_spec(spec, device=device, stream=stream.cuda_stream)
        b = ctx.batch(ps)
        ts = []
        for it in range(3):
            ctx.set_x(x0)
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            b.solve(None, MAXITERS, FTOL)
            e.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        r = b.fetch()
        row = {"nv": int(len(ps.vids)), "nf": int(len(ps.fids)), "gpu_ms_ssmaxit25": float(min(ts[1:])), "f_init": float(r["f_init"][0]),
               "f_end": float(r["f_end"][0]), "evaluations": int(r["n_feval"][0]), "us_per_evaluation": float(min(ts[1:])) * 1e3 / max(int(r["n_feval"][0]), 1),
               "mapping": b.info(), "kernel": "solve_grid_kernel<BaOps> (cooperative grid)"}
        if with_cpu:
            from oracle import oracle_py as O
            xs = x0[ps.vids]
            od = O.OracleFunction.from_spec(spec, "devtrig")
            od.set_x(x0)
            o = od.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, xs, 1, FTOL)
            ctx.set_x(x0)
            r1 = ctx.solve_cgd(ps, xs, 1, FTOL)
            sc = Context.from_spec(spec, device=device, stream=stream.cuda_stream)
            sc.set_option("strict", 1)
            sc.set_x(x0)
            t0 = time.perf_counter()
            rs = sc.solve_cgd(ps, xs, 1, FTOL)
            ts_strict = time.perf_counter() - t0
            row["parity_ssmaxit1"] = {"cpu_oracle_s": o["seconds"], "f_end_cpu_devtrig": float(o["f_end"][0]), "f_end_gpu": float(r1["f_end"][0]),
                                      "rel_diff_production": float(abs(r1["f_end"][0] - o["f_end"][0]) / abs(o["f_end"][0])),
                                      "strict_bit_identical": bool(rs["f_end"][0] == o["f_end"][0] and np.array_equal(rs["x"], o["x"])),
                                      "strict_host_call_ms": ts_strict * 1e3, "evaluations": int(r1["n_feval"][0]),
                                      "cpu_seconds_per_evaluation": o["seconds"] / max(int(r1["n_feval"][0]), 1)}
            row["estimated_cpu_s_ssmaxit25"] = o["seconds"] / max(int(r1["n_feval"][0]), 1) * int(r["n_feval"][0])
        res[name] = row
        b.close()
        del ctx
    return res


def cfg4_sibling_wave(device, stream, rank, world, dist, dev):
    """Synthetic factor graph of BASELINE config 4 (sinusoid tree h=19 k=2 arity 4: 1,048,575 variables, 4,194,292
    NonlinearProductFactors).  With the top 10 tree levels assigned the graph falls apart into 1024 sibling
    components (1023 variables / 4092 factors each); they are dealt to the ranks (LPT shard, static CSR replicated),
    every rank solves its share as ONE batch, the objective is all-reduced.  STRONG scaling: total work fixed."""
    import torch
    from rdis_b200 import Context, problems as P
    spec = P.sinusoid(19, 2, 4)
    x0 = P.random_start(spec, 1)
    ps = P.sinusoid_subtree_problems(spec, 10)
    mine = ps.subset(P.shard_problems(ps, rank, world)) if world > 1 else ps
    ctx = Context.from_spec(spec, device=device, stream=stream.cuda_stream)
    x0_dev = torch.from_numpy(x0).to(dev)
    obj = torch.zeros(1, dtype=torch.float64, device=dev)
    b = ctx.batch(mine)
    times = []
    for it in range(3):
        ctx.set_x_device(x0_dev.data_ptr(), spec["V"])
        obj.zero_()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        b.solve(None, MAXITERS, FTOL)
        b.objective_device(obj.data_ptr())
        if dist is not None:
            dist.all_reduce(obj)
        e.record(stream)
        torch.cuda.synchronize()
        times.append(a.elapsed_time(e))
    ms = float(np.min(times[1:]))
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    r = b.fetch()
    hist = np.bincount(r["status"], minlength=8).astype(np.int64)
    if dist is not None:
        t = torch.from_numpy(hist).to(dev)
        dist.all_reduce(t)
        hist = t.cpu().numpy()
    out = {"workload": "sinusoid h=19 k=2 arity=4 (V=%d, F=%d): %d sibling components of %d variables / %d factors after assigning "
                       "the top 10 tree levels, %d per rank" % (spec["V"], spec["F"], ps.n, ps.var_off[1], ps.fac_off[1], mine.n),
           "scaling": "strong", "n_gpus": world, "ms": ms, "solves_per_sec": ps.n / (ms * 1e-3),
           "objective_sum_f_end": float(obj.item()), "mapping": b.info(),
           "status_histogram": hist.tolist(), "status_names": "ftol, gtol, gg_zero, maxiters, dbrent_itmax, empty, nonfinite (safety exit), bracket_cap (safety exit)",
           "kernel": "solve_nlpf_resident_kernel (one CTA per component, resident in shared memory)",
           "dram_bytes_per_launch_ncu": measured_traffic("solve_nlpf_resident_kernel"),
           "fp64_pipe_active_pct_ncu": measured_traffic("solve_nlpf_resident_kernel", "fp64_pipe_active_pct")}
    if rank == 0:
        # the callers either side of the solve on the same graph (host buffers in and out, wall clock around the C call):
        # sibling membership (rdisgpu_components) and interval bounds of every factor (rdisgpu_bounds)
        assigned = np.zeros(spec["V"], np.uint8); assigned[:1023] = 1     # the top 10 tree levels
        ctx.set_x(x0)
        ctx.components(assigned)
        t0 = time.perf_counter(); vl, fl, ncomp, rounds = ctx.components(assigned); t1 = time.perf_counter()
        out["membership"] = {"what": "rdisgpu_components: labels of %d variables / %d factors, host buffers" % (spec["V"], spec["F"]),
                             "ms": (t1 - t0) * 1e3, "components": int(ncomp), "rounds": int(rounds),
                             "matches_generator": bool(ncomp == ps.n and np.array_equal(np.sort(np.nonzero(vl == vl[ps.vids[0]])[0]),
                                                                                         np.sort(ps.vids[:ps.var_off[1]])))}
        ctx.bounds(assigned)
        t0 = time.perf_counter(); lo, hi, tot = ctx.bounds(assigned); t1 = time.perf_counter()
        out["bounds"] = {"what": "rdisgpu_bounds: interval bounds of all %d factors (top 10 levels assigned, the rest at their domains), "
                                 "host buffers, per-factor bounds returned" % spec["F"],
                         "ms": (t1 - t0) * 1e3, "factor_bounds_per_sec": spec["F"] / (t1 - t0), "sum": [tot[0], tot[1]],
                         "parity": "unpinned upstream (Boost.Interval not vendored); tested against oracle/interval_oracle.hpp"}
    return out


def sweep_roofline(device, stream):
    """Residual sweep (evalFactors over all factors, every factor's value written back) and the
    gradient sweep on the cfg4 sinusoid graph.  268.4 MB of algorithmic bytes per eval launch
    (20 F + 21 E + 8 V) — larger than L2, and L2 is flushed between launches anyway.
    Timed with CUDA events around the asynchronous device-pointer entry point (rdisgpu_eval_device):
    nothing but the kernel is inside the events."""
    import torch
    from rdis_b200 import Context, problems as P
    spec = P.sinusoid(19, 2, 4)
    V, F, E = spec["V"], spec["F"], len(spec["vid"])
    ctx = Context.from_spec(spec, device=device, stream=stream.cuda_stream)
    ctx.set_x(P.random_start(spec, 1))
    dev = torch.device("cuda", device)
    per_factor = torch.empty(F, dtype=torch.float64, device=dev)
    total = torch.zeros(1, dtype=torch.float64, device=dev)
    grad = torch.empty(V, dtype=torch.float64, device=dev)
    flush = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    peak, peak_src = peaks()

    def timed(fn, reps=12):
        """(a) one launch per event pair, L2 flushed before each; (b) `reps` back-to-back launches inside one
        event pair (inputs are 2.1x the L2, the streamed slices are loaded evict-first): average per launch."""
        for _ in range(3):
            fn()
        times = []
        for _ in range(reps):
            flush.zero_()   # evict ...
            flush.sum()     # ... then a read pass, so that no dirty lines are written back inside the timing
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        flush.zero_()
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return float(np.mean(times)), float(np.min(times)), a.elapsed_time(b) / reps

    ms_flushed, ms_min, ms = timed(lambda: ctx.eval_device(total.data_ptr(), per_factor.data_ptr()))
    s = float(total.item())
    assert abs(float(per_factor.sum().item()) - s) <= 1e-9 * abs(s)
    abytes = 20.0 * F + 21.0 * E + 8.0 * V
    ach = abytes / (ms * 1e-3) / 1e9
    out = {"bound": "hbm", "kernel": "nlpf_tile_sweep_kernel<false>", "workload": "sinusoid h=19 k=2 arity=4: V=%d F=%d E=%d" % (V, F, E),
           "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
           "traffic": measured_traffic("nlpf_tile_sweep_kernel<false>"),
           "algorithmic_bytes_per_launch": abytes, "launch_ms": ms, "launch_ms_single_flushed": ms_flushed,
           "frac_single_flushed": abytes / (ms_flushed * 1e-3) / 1e9 / peak, "launch_ms_min": ms_min,
           "factor_evals_per_sec": F / (ms * 1e-3), "sum": s, "peak_source": peak_src,
           "timing": "launch_ms = 12 back-to-back launches inside one CUDA-event pair / 12 (inputs 268 MB > 126 MB L2, streams "
                     "loaded evict-first); launch_ms_single_flushed = one launch per event pair, L2 flushed (256 MiB memset + read "
                     "pass, untimed) before each: includes ~3 us of launch latency"}
    # eval + gradient: tile kernel writes the per-edge partials, the variable-major gather folds them
    msg_flushed, msg_min, msg = timed(lambda: ctx.grad_device(grad.data_ptr()))
    gbytes = 20.0 * F + 21.0 * E + 16.0 * V
    out["grad_sweep"] = {"kernels": "nlpf_tile_sweep_kernel<true> + gather_grad_kernel<NlpfOps>", "launch_ms": msg, "launch_ms_single_flushed": msg_flushed, "launch_ms_min": msg_min,
                         "algorithmic_bytes": gbytes, "achieved": gbytes / (msg * 1e-3) / 1e9, "frac": gbytes / (msg * 1e-3) / 1e9 / peak,
                         "unit": "GB/s", "grad_norm": float(grad.norm().item())}
    return out


def parity_leg(spec, pts, cams, device, stream):
    """The step's results against the CPU oracle on the SAME inputs (the real graph, the file's state), production
    kernels: the point wave against the devtrig twin (bit identity expected) and the glibc oracle (1e-6), the full
    step objective of the strict mode against the devtrig twin's sequence (bit identity expected).  The camera wave's
    per-component comparison lives in profiles/r02_parity_gpu.json (tools/parity_gpu.py: minutes of CPU)."""
    from rdis_b200 import Context
    from oracle import oracle_py as O
    x0 = spec["x0"]
    ctx = Context.from_spec(spec, device=device, stream=stream.cuda_stream)
    ctx.set_x(x0)
    f0 = ctx.eval()
    rp = ctx.solve_cgd(pts, x0[pts.vids], MAXITERS, FTOL)
    x1 = ctx.get_x()
    rc = ctx.solve_cgd(cams, x1[cams.vids], MAXITERS, FTOL)
    out = {"graph": "data/ladybug-problem-49-7776-pre.txt (tests/golden/ladybug_49_7776.npz), x0 = file state", "objective_start": f0,
           "objective_after_points_then_cameras_production": float(rc["f_end"].sum())}
    for variant, label in (("devtrig", "oracle_devtrig"), ("restated", "oracle_glibc")):
        orc = O.OracleFunction.from_spec(spec, variant)
        orc.set_x(x0)
        op = orc.solve_cgd_batch(pts.var_off, pts.vids, pts.fac_off, pts.fids, x0[pts.vids], MAXITERS, FTOL)
        rel = np.abs(rp["f_end"] - op["f_end"]) / np.maximum(np.abs(op["f_end"]), 1e-300)
        out["point_wave_vs_" + label] = {"components": int(pts.n), "bit_identical": int((rp["f_end"] == op["f_end"]).sum()),
                                         "within_1e-6": int((rel <= 1e-6).sum()), "max_rel": float(rel.max()),
                                         "rel_diff_of_sums": float(abs(rp["f_end"].sum() - op["f_end"].sum()) / abs(op["f_end"].sum())),
                                         "cpu_oracle_s": op["seconds"]}
    return out


def cpu_baseline(spec, pts, cams, x0, budget_s):
    """The CPU oracle (single-threaded reference-semantics port) on a bounded sample of the same
    wave: every k-th point component and a few camera components, scaled to the whole wave."""
    from oracle import oracle_py as O
    O.build()
    orc = O.OracleFunction.from_spec(spec)
    orc.set_x(x0)
    # calibrate on a handful, then size the sample for ~budget_s
    probe = pts.subset(range(0, pts.n, max(1, pts.n // 32)))
    t = orc.solve_cgd_batch(probe.var_off, probe.vids, probe.fac_off, probe.fids, x0[probe.vids], MAXITERS, FTOL)["seconds"]
    per_pt = t / probe.n
    cam_probe = cams.subset([0])
    orc.set_x(x0)
    t_cam = orc.solve_cgd_batch(cam_probe.var_off, cam_probe.vids, cam_probe.fac_off, cam_probe.fids, x0[cam_probe.vids],
                                MAXITERS, FTOL)["seconds"]
    n_cam = int(max(1, min(cams.n, (0.5 * budget_s) // max(t_cam, 1e-3))))
    n_pt = int(max(32, min(pts.n, (0.5 * budget_s) // max(per_pt, 1e-6))))
    ps_pt = pts.subset(np.linspace(0, pts.n - 1, n_pt).astype(int))
    ps_cam = cams.subset(np.linspace(0, cams.n - 1, n_cam).astype(int))
    orc.set_x(x0)
    orc.reset_counters()
    r_pt = orc.solve_cgd_batch(ps_pt.var_off, ps_pt.vids, ps_pt.fac_off, ps_pt.fids, x0[ps_pt.vids], MAXITERS, FTOL)
    orc.set_x(x0)
    r_cam = orc.solve_cgd_batch(ps_cam.var_off, ps_cam.vids, ps_cam.fac_off, ps_cam.fids, x0[ps_cam.vids], MAXITERS, FTOL)
    cnt = orc.counters()
    wave_s = r_pt["seconds"] * pts.n / ps_pt.n + r_cam["seconds"] * cams.n / ps_cam.n
    return {"value": (pts.n + cams.n) / wave_s, "unit": "solves/s", "cores": 1, "kind": "port",
            "sample": "%d of %d point components + %d of %d camera components, scaled to the %d-solve wave "
                      "(estimated CPU wave time %.1f s)" % (ps_pt.n, pts.n, ps_cam.n, cams.n, pts.n + cams.n, wave_s),
            "point_solves_per_sec": ps_pt.n / r_pt["seconds"], "camera_solves_per_sec": ps_cam.n / r_cam["seconds"],
            "factor_eval_calls_per_sec": cnt["factor_eval_calls"] / (r_pt["seconds"] + r_cam["seconds"]),
            "host_cpus": os.cpu_count(), "oracle": "oracle/liboracle.so (-O2, 1 thread)"}


# ------------------------------------------------------------------------------------------
# reference arm: the CPU implementation of the path on all host threads
# ------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    problems = load_problems_module()  # numpy only: librdis_b200.so is never mapped on this arm
    from oracle import oracle_py as O
    O.build()
    variant = "refnrc" if O.have_refnrc() else "restated"
    spec = problems.load_golden_ba()
    x0 = spec["x0"]
    pts, cams = problems.ba_point_problems(spec), problems.ba_camera_problems(spec)
    nthreads = max(1, os.cpu_count() or 1)
    reps = [O.OracleFunction.from_spec(spec, variant) for _ in range(nthreads)]
    for r in reps:
        r.set_x(x0)
    # bounded sample per step: sized from a probe so that (steps+warmup) steps end within minutes
    probe = pts.subset(range(0, pts.n, max(1, pts.n // 64)))
    tp = reps[0].solve_cgd_batch(probe.var_off, probe.vids, probe.fac_off, probe.fids, x0[probe.vids], MAXITERS, FTOL)["seconds"] / probe.n
    cprobe = cams.subset([0])
    reps[0].set_x(x0)
    tc = reps[0].solve_cgd_batch(cprobe.var_off, cprobe.vids, cprobe.fac_off, cprobe.fids, x0[cprobe.vids], MAXITERS, FTOL)["seconds"]
    total_steps = args.steps + args.warmup
    budget = min(8.0, 150.0 / max(total_steps, 1))          # seconds of wall clock per step
    n_cam = int(max(nthreads, min(cams.n, (0.5 * budget * nthreads) // max(tc, 1e-3))))
    n_cam = min(cams.n, n_cam)
    n_pt = int(max(nthreads * 8, min(pts.n, (0.5 * budget * nthreads) // max(tp, 1e-6))))
    n_pt = min(pts.n, n_pt)
    ps_pt = pts.subset(np.linspace(0, pts.n - 1, n_pt).astype(int))
    ps_cam = cams.subset(np.linspace(0, cams.n - 1, n_cam).astype(int))

    def step():
        for r in reps:
            r.set_x(x0)
        a = reps[0].solve_cgd_batch(ps_pt.var_off, ps_pt.vids, ps_pt.fac_off, ps_pt.fids, x0[ps_pt.vids], MAXITERS, FTOL, replicas=reps)
        b = reps[0].solve_cgd_batch(ps_cam.var_off, ps_cam.vids, ps_cam.fac_off, ps_cam.fids, x0[ps_cam.vids], MAXITERS, FTOL, replicas=reps)
        # whole-wave time at this throughput
        return a["seconds"] * pts.n / ps_pt.n + b["seconds"] * cams.n / ps_cam.n, a["seconds"] + b["seconds"]

    for _ in range(args.warmup):
        step()
    wave, spent = [], 0.0
    for _ in range(args.steps):
        w, s = step()
        wave.append(w)
        spent += s
    wave_s = float(np.mean(wave))
    value = (pts.n + cams.n) / wave_s
    sample = ("each step: %d of %d point + %d of %d camera components on %d threads (one function replica per thread), "
              "scaled to the 7825-solve wave; the total work is the same at every N (strong scaling)" % (ps_pt.n, pts.n, ps_cam.n, cams.n, nthreads))
    out = {"impl": "reference", "metric": "subspace-solves/sec", "value": value, "unit": "solves/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": wave_s * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": DATA,
           "config": config(args.gpus),
           "cpu_baseline": {"value": value, "unit": "solves/s", "cores": nthreads,
                            "kind": "port", "sample": sample,
                            "driver": "reference minimize_nrc.h (oracle/_ref)" if variant == "refnrc" else "restated NR driver"},
           "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "cpu_seconds_measured": spent}
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
