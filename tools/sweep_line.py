"""One-line summary of tools/sweep_probe.py's JSON (value sweep and gradient sweep): python tools/sweep_probe.py | python tools/sweep_line.py"""
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); g=d.get("grad_sweep")
print("eval %.1f us frac %.3f | grad %.1f us (flushed %.1f) frac %.3f norm %r" % (d["launch_ms"]*1e3, d["frac"], g["launch_ms"]*1e3, g["launch_ms_single_flushed"]*1e3, g["frac"], g["grad_norm"]))
