"""Prints the solve-kernel times of the bench step for the library named by RDIS_B200_LIB (kernel tuning only)."""
import json, subprocess, sys, os
out = subprocess.run([sys.executable, "bench.py", "--no-sweep", "--no-cpu", "--steps", "30"], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    m = d["mapping"]["cameras"]
    print(os.environ.get("RDIS_B200_LIB", "default").split("/")[-1], "step %.3f ms" % d["ms_per_step"], {k: round(v, 3) for k, v in d["kernel_ms"].items()},
          "C=%d T=%d" % (m["cluster_size"], m["camera_threads"]), "e2e %.0f" % d["e2e"]["value"], "obj %.6f" % d["objective_after_step"])
except Exception as e:
    print("failed", e, out.stderr[-500:])
