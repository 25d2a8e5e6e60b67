"""Runs only the plugin end-to-end leg of bench.py (tests/native/host_driver benchwaves on the real graph).
usage: python tools/e2e_probe.py [steps]"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    steps = sys.argv[1] if len(sys.argv) > 1 else "10"
    from rdis_b200 import problems
    spec = problems.load_golden_ba()
    with tempfile.TemporaryDirectory() as td:
        bal = os.path.join(td, "ladybug.txt")
        bench.write_bal(spec, bal)
        for _ in range(2):
            p = subprocess.run([os.path.join(ROOT, "tests", "native", "host_driver"), "benchwaves", bal, steps, "3"],
                               capture_output=True, text=True)
            print(p.stdout.strip().splitlines()[-1] if p.returncode == 0 else p.stderr[-400:])
            if p.stderr.strip():
                print("\n".join(p.stderr.strip().splitlines()[-12:]))
