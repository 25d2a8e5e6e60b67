"""CPU tests of the host side: the C-ABI library loads and exports what include/rdis_gpu.h
declares (no compute without a GPU), the problem generators agree with the oracle's line-by-line
restatements of the reference's builders, and the component sharding used at N>1 is exact."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "rdis_gpu.h")).read()
    return sorted(set(re.findall(r"RDISGPU_API\s+[\w\s\*]+?\b(rdisgpu_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built_lib):
    from rdis_b200 import capi
    names = _declared_symbols()
    assert len(names) >= 25
    lib = C.CDLL(built_lib)
    for n in names:
        assert hasattr(lib, n), "librdis_b200.so does not export " + n
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS out of date with include/rdis_gpu.h"
    lib.rdisgpu_version.restype = C.c_char_p
    assert b"sm_100a" in lib.rdisgpu_version()


def test_no_cpu_fallback(built_lib):
    """Without a usable sm_100 device the library refuses to create a context (and says why);
    there is nothing else it could run on."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rdis_b200 import Context, RdisGpuError
    with pytest.raises(RdisGpuError) as e:
        Context(0)
    assert "CUDA" in str(e.value) or "device" in str(e.value)


def test_product_package_does_not_touch_the_oracle():
    """Nothing under rdis_b200/ may import, include, link or execute oracle/ (comments that cite the
    tests which compare against it are fine)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rdis_b200")):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                for line in open(path):
                    code = line.split("#")[0]
                    assert not re.search(r"\b(import|from)\s+oracle\b|liboracle|oracle_py", code), (path, line)
            elif f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                for line in open(path, errors="ignore"):
                    code = line.split("//")[0]
                    if "#include" in code or "-l" in code or "-I" in code:
                        assert "oracle" not in code, (path, line)


def test_sinusoid_generator_matches_reference_restatement(oracle_mod):
    """rdis_b200.problems.sinusoid (vectorised) == makeHighDimSinusoid as restated line by line in the
    oracle (src/OptimizableFunctionGenerator.cpp:660-760): same factor order, edges, coefficients."""
    from rdis_b200 import problems as P
    for h, k, ar, odd in [(6, 3, 3, False), (5, 2, 4, True), (9, 1, 3, False), (1, 9, 2, False), (3, 2, 8, True)]:
        a = P.sinusoid(h, k, ar, odd)
        b = oracle_mod.OracleFunction.sinusoid(h, k, ar, odd).export()
        assert a["V"] == b["V"] and a["F"] == b["F"], (h, k, ar)
        for key in ("rowptr", "vid", "expo", "konst", "sine", "coeff", "lb", "ub", "samp_lo", "samp_hi"):
            assert np.array_equal(np.asarray(a[key], dtype=b[key].dtype), b[key]), (h, k, ar, key)


def test_config_sizes_of_survey_table():
    from rdis_b200 import problems as P
    s = P.sinusoid(999, 1, 3)
    assert (s["V"], s["F"], len(s["vid"])) == (1000, 2999, 3998)          # cfg2 chain
    s = P.sinusoid(6, 3, 3)
    assert (s["V"], s["F"], len(s["vid"])) == (1093, 3278, 4370)          # cfg2 tree


def test_ba_domains_match_reference_restatement(oracle_mod):
    """problems.ba_domains == BundleAdjustmentFunction::setDomain as restated in the oracle's BAL
    loader (the fixture's lb/ub were written by that loader from the reference's data file)."""
    from rdis_b200 import problems as P
    z = np.load(os.path.join(P.GOLDEN_DIR, "ladybug_49_7776.npz"))
    lb, ub, _, _ = P.ba_domains(z["x0"], 49)
    assert np.array_equal(lb, z["lb"]) and np.array_equal(ub, z["ub"])


def test_problem_sets_are_sibling_components():
    """Point / camera problem sets: disjoint variables and factors, every factor of a problem touches
    the problem's own block (the precondition of the sibling batch, src/Component.cpp:508-549)."""
    from rdis_b200 import problems as P
    spec = P.ba_synthetic(ncams=9, npts=200, nobs=800, seed=1)
    pts, cams = P.ba_point_problems(spec), P.ba_camera_problems(spec)
    assert pts.n == 200 and cams.n == 9
    assert len(np.unique(pts.fids)) == spec["F"] == len(np.unique(cams.fids))
    for i in range(pts.n):
        f = pts.fids[pts.fac_off[i]:pts.fac_off[i + 1]]
        assert (spec["pt"][f] == i).all() and (np.diff(f) > 0).all()
        assert np.array_equal(pts.vids[3 * i:3 * i + 3], 9 * 9 + 3 * i + np.arange(3))
    for c in range(cams.n):
        f = cams.fids[cams.fac_off[c]:cams.fac_off[c + 1]]
        assert (spec["cam"][f] == c).all() and (np.diff(f) > 0).all()
    sn = P.sinusoid(6, 2, 4)
    sub = P.sinusoid_subtree_problems(sn, 3)
    assert sub.n == 8
    assert len(np.unique(sub.vids)) == len(sub.vids) and len(np.unique(sub.fids)) == len(sub.fids)
    # every factor of a subtree problem has all its variables inside the subtree or among the assigned ancestors
    depth_assigned = set(range(7))  # BFS ids of depth < 3 in a binary tree
    for i in range(sub.n):
        own = set(sub.vids[sub.var_off[i]:sub.var_off[i + 1]].tolist())
        for f in sub.fids[sub.fac_off[i]:sub.fac_off[i + 1]]:
            vs = sn["vid"][sn["rowptr"][f]:sn["rowptr"][f + 1]].tolist()
            assert all(v in own or v in depth_assigned for v in vs) and any(v in own for v in vs)


def test_shard_problems_partition_is_exact_and_balanced():
    from rdis_b200 import problems as P
    spec = P.ba_synthetic(ncams=12, npts=500, nobs=2100, seed=4)
    for ps in (P.ba_point_problems(spec), P.ba_camera_problems(spec)):
        for world in (1, 2, 4, 8):
            parts = [P.shard_problems(ps, r, world) for r in range(world)]
            allidx = np.concatenate(parts)
            assert np.array_equal(np.sort(allidx), np.arange(ps.n))          # every component exactly once
            cost = np.diff(ps.fac_off) * np.diff(ps.var_off)
            loads = np.array([cost[p].sum() for p in parts], dtype=float)
            assert loads.max() <= loads.mean() + cost.max()                  # LPT bound


@pytest.mark.timeout(300)
def test_two_rank_component_shard_gloo(tmp_path):
    """world_size-2 gloo run of the N>1 host logic: each rank owns the shard shard_problems gives it,
    solves it (CPU oracle standing in for the device here — this test is about the partition and the
    collective, not the arithmetic), and the all-reduced objective equals the single-process total."""
    script = tmp_path / "two_rank.py"
    script.write_text('''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from rdis_b200 import problems as P
from rdis_b200.shard import allreduce_objective, gather_solution
from oracle import oracle_py as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
spec = P.ba_synthetic(ncams=4, npts=40, nobs=150, seed=8)
x0 = spec["x0"]
ps = P.ba_point_problems(spec)
mine = P.shard_problems(ps, rank, world)
sub = ps.subset(mine)
orc = O.OracleFunction.from_spec(spec); orc.set_x(x0)
o = orc.solve_cgd_batch(sub.var_off, sub.vids, sub.fac_off, sub.fids, x0[sub.vids], 25, 3e-8)
total = allreduce_objective(torch.tensor([o["f_end"].sum()], dtype=torch.float64))
x = gather_solution(torch.from_numpy(x0.copy()), torch.from_numpy(sub.vids.astype(np.int64)), torch.from_numpy(o["x"]))
if rank == 0:
    orc1 = O.OracleFunction.from_spec(spec); orc1.set_x(x0)
    full = orc1.solve_cgd_batch(ps.var_off, ps.vids, ps.fac_off, ps.fids, x0[ps.vids], 25, 3e-8)
    assert abs(float(total) - full["f_end"].sum()) <= 1e-9 * abs(full["f_end"].sum()), (float(total), full["f_end"].sum())
    xs = x0.copy(); xs[ps.vids] = full["x"]
    assert np.array_equal(x.numpy(), xs)
    print("TWO_RANK_OK")
dist.destroy_process_group()
''' % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=280)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "TWO_RANK_OK" in out.stdout


@pytest.mark.timeout(300)
def test_two_rank_strong_shard_step_gloo(tmp_path):
    """world_size-2 gloo run of bench.py's multi-GPU step logic (the CPU oracle standing in for the device): ONE
    sibling set dealt to the ranks by bench.lpt_shard, the point wave's results exchanged with the padded equal-slice
    all-gather the bench uses and scattered into every replica, the camera wave solved on the gathered state, the
    objective all-reduced — equal to the single-process alternating step to the bit (the solves are independent and
    deterministic; only the partition differs)."""
    script = tmp_path / "strong_shard.py"
    script.write_text('''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import bench
from rdis_b200 import problems as P
from oracle import oracle_py as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
spec = P.ba_synthetic(ncams=5, npts=70, nobs=300, seed=4)
x0 = spec["x0"]
pts, cams = P.ba_point_problems(spec), P.ba_camera_problems(spec)
own_p, own_c = bench.lpt_shard(pts, world), bench.lpt_shard(cams, world)
assert np.array_equal(np.sort(np.concatenate(own_p)), np.arange(pts.n)) and np.array_equal(np.sort(np.concatenate(own_c)), np.arange(cams.n))
nf = np.diff(cams.fac_off); loads = [nf[o].sum() for o in own_c]
assert max(loads) <= np.mean(loads) + nf.max()                       # LPT bound
my_pts, my_cams = pts.subset(own_p[rank]), cams.subset(own_c[rank])
orc = O.OracleFunction.from_spec(spec); orc.set_x(x0)
a = orc.solve_cgd_batch(my_pts.var_off, my_pts.vids, my_pts.fac_off, my_pts.fids, x0[my_pts.vids], 25, 3e-8)
# the bench's exchange: equal-sized slices, the pad repeats the first (vid, value) pair of the rank
nmax = max(len(pts.subset(o).vids) for o in own_p)
pad = lambda v: np.concatenate([v, np.full(nmax - len(v), v[0], v.dtype)])
vid_all = np.concatenate([pad(pts.subset(o).vids) for o in own_p])
send = torch.from_numpy(pad(a["x"]))
recv = torch.zeros(world * nmax, dtype=torch.float64)
dist.all_gather_into_tensor(recv, send)
state = x0.copy(); state[vid_all] = recv.numpy()
orc.set_x(state)
b = orc.solve_cgd_batch(my_cams.var_off, my_cams.vids, my_cams.fac_off, my_cams.fids, state[my_cams.vids], 25, 3e-8)
part = torch.tensor([b["f_end"].sum()], dtype=torch.float64)
dist.all_reduce(part)
if rank == 0:
    one = O.OracleFunction.from_spec(spec); one.set_x(x0)
    pa = one.solve_cgd_batch(pts.var_off, pts.vids, pts.fac_off, pts.fids, x0[pts.vids], 25, 3e-8)
    full = x0.copy(); full[pts.vids] = pa["x"]
    assert np.array_equal(state, full)
    one.set_x(full)
    ca = one.solve_cgd_batch(cams.var_off, cams.vids, cams.fac_off, cams.fids, full[cams.vids], 25, 3e-8)
    want = sum(ca["f_end"][own_c[r]].sum() for r in range(world))   # the all-reduce adds rank partials
    assert float(part) == want, (float(part), want)
    assert abs(float(part) - ca["f_end"].sum()) <= 1e-12 * abs(ca["f_end"].sum())
    print("STRONG_SHARD_OK")
dist.destroy_process_group()
''' % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29537", str(script)],
                         capture_output=True, text=True, timeout=280)
    assert out.returncode == 0, out.stderr[-3000:] + out.stdout[-1000:]
    assert "STRONG_SHARD_OK" in out.stdout


def test_synthetic_ba_generator_rejects_impossible_sizes():
    """Every point of the synthetic bundle-adjustment graph is seen by 2..min(29, ncams) DISTINCT cameras: an observation
    count outside [2 npts, min(29, ncams) npts] has no such graph (the generator used to loop forever on it)."""
    from rdis_b200 import problems as P
    with pytest.raises(ValueError):
        P.ba_synthetic(ncams=6, npts=700, nobs=5000, seed=9)
    with pytest.raises(ValueError):
        P.ba_synthetic(ncams=6, npts=700, nobs=1000, seed=9)
    spec = P.ba_synthetic(ncams=6, npts=700, nobs=4200, seed=9)  # the densest graph there is: every point in every camera
    assert spec["F"] == 4200 and int(np.bincount(spec["pt"]).min()) == 6


def test_bench_reads_the_committed_ncu_numbers():
    """bench.py's roofline.traffic / fp64-pipe figures come from profiles/r02_traffic.json (tools/make_traffic.py over the
    committed ncu summaries): the kernels bench.py asks for must be there under the names it uses."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for kernel in ("solve_ba_cameras_kernel", "solve_ba_points_kernel", "nlpf_tile_sweep_kernel<false>", "solve_nlpf_resident_kernel"):
        dram = bench.measured_traffic(kernel)
        assert dram is not None and dram > 0, kernel
        assert bench.measured_traffic(kernel, "fp64_pipe_active_pct") is not None, kernel
    # the streaming sweep moves about its algorithmic bytes, the latency-bound solve kernels almost nothing
    assert 0.8 * 268.4e6 < bench.measured_traffic("nlpf_tile_sweep_kernel<false>") < 1.2 * 268.4e6
    assert bench.measured_traffic("solve_ba_cameras_kernel") < 16e6
