"""Runs only the cfg4 sweep leg of bench.py (for ncu captures and quick timing).
usage: python tools/sweep_probe.py [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

if __name__ == "__main__":
    torch.cuda.set_device(0)
    out = bench.sweep_roofline(0, torch.cuda.current_stream())
    print(json.dumps(out))
