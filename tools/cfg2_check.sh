python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "config2" 2>&1 | grep "config 2\|passed\|failed\|Error\|assert" | head
python -c "
import json, torch, bench
torch.cuda.set_device(0)
print(json.dumps(bench.cfg2_full_solve(0, torch.cuda.current_stream())))"
