"""profiles/r02_ncu_*_raw.txt (tools/ncu_raw.py summaries) -> profiles/r02_traffic.json: per kernel the DRAM bytes per launch,
fp64-pipe %, issue %, DMMA / tensor-pipe %, registers and duration under ncu, averaged over the captured launches.
bench.py reads it into roofline.traffic.  usage: python tools/make_traffic.py"""
import glob, json, os, re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def short(name):
    name = name.replace("void ", "").replace("rdisgpu::", "").strip()
    m = re.match(r"([A-Za-z0-9_]+)(<[^>(]*>)?", name)
    base, targs = m.group(1), (m.group(2) or "")
    if base == "ba_sweep_kernel":
        return "ba_sweep_kernel<smem,%s>" % ("rows" if targs.replace(" ", "") in ("<1,1>", "<(bool)1,(bool)1>") else "values")
    if base in ("solve_ba_points_kernel", "solve_ba_cameras_kernel", "lm_potrf_kernel", "cc_hook_kernel", "cc_jump_kernel", "lm_trsv_kernel"):
        return base
    return base + targs


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_ncu_*_raw.txt"))):
        cur = None
        for line in open(path):
            parts = line.split()
            if line.strip().startswith("Kernel Name"):
                cur = out.setdefault(short(line.split("Kernel Name", 1)[1]), {"launches_captured": 0, "src": os.path.basename(path), "_acc": {}})
                cur["launches_captured"] += 1
                cur["src"] = os.path.basename(path)
                continue
            if cur is None or len(parts) < 2:
                continue
            metric = parts[0]
            try:
                val = float(parts[1])
            except ValueError:
                continue
            unit = parts[2] if len(parts) > 2 else ""
            key = {"dram__bytes_read.sum": "dram_bytes", "dram__bytes_write.sum": "dram_bytes", "gpu__time_duration.sum": "us_per_launch_under_ncu",
                   "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "registers",
                   "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
                   "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
                   "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active": "dmma_issue_pct_of_peak",
                   "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct"}.get(metric)
            if key is None:
                continue
            if key in ("dram_bytes", "us_per_launch_under_ncu"):
                val *= UNIT.get(unit, 1.0)
            cur["_acc"][key] = cur["_acc"].get(key, 0.0) + val
    for k, v in out.items():
        n = v["launches_captured"]
        for key, tot in v.pop("_acc").items():
            v[key] = tot / n
    # aliases bench.py asks for: <false> / <true> spellings and the bare name of a kernel with one captured instantiation
    for k in list(out.keys()):
        for a, b in (("<0>", "<false>"), ("<1>", "<true>")):
            if k.endswith(a):
                out[k[:-len(a)] + b] = out[k]
        base = k.split("<")[0]
        if base != k and sum(1 for kk in out if kk.split("<")[0] == base and not kk.endswith(("<false>", "<true>"))) == 1:
            out.setdefault(base, out[k])
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print("%-44s x%-3d %10.1f us  %12.0f B DRAM  fp64 %5.1f %%  issue %5.1f %%" % (k, v["launches_captured"], v.get("us_per_launch_under_ncu", 0), v.get("dram_bytes", 0), v.get("fp64_pipe_active_pct", 0), v.get("issue_active_pct", 0)))


if __name__ == "__main__":
    main()
