#!/bin/bash
# End-of-round evidence (run under gpurun, ONE GPU): launch list of the bench step, ncu --set full of the solve kernels and of
# the gradient-sweep kernels (summarised on the box), then both bench arms.  Copy gpurun_out/r02_* and bench_*.json to profiles/.
NCU="ncu --set full --clock-control none --import-source on"
summarise() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw.py > gpurun_out/r02_ncu_$1_raw.txt 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_stalls.py --flow > gpurun_out/r02_ncu_$1_stalls.txt 2>&1
  rm -f gpurun_out/$1.ncu-rep
}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep > /dev/null 2>&1
$NCU -k regex:"solve_ba_(cameras|points)" -s 6 -c 2 -o gpurun_out/solve_ba python bench.py --steps 1 --warmup 3 --no-cpu --no-sweep > /dev/null 2>&1; summarise solve_ba
$NCU -k regex:"nlpf_tile_sweep|gather_grad" -s 26 -c 3 -o gpurun_out/nlpf_grad python tools/sweep_probe.py > /dev/null 2>&1; summarise nlpf_grad
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ls -la gpurun_out | head -30
