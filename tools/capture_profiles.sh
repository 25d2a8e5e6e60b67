#!/bin/bash
# Round-2 evidence capture (run under gpurun, ONE GPU): launch list of the bench step + ncu --set full of every kernel family.
# Reports land in gpurun_out/r02_*.ncu-rep; tools/summarise_profiles.sh turns them into profiles/r02_*.txt here.
set -x
NCU="ncu --set full --clock-control none --import-source on"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep > /dev/null 2>&1
$NCU -k regex:"solve_ba_(cameras|points)" -s 6 -c 2 -o gpurun_out/r02_solve_ba python bench.py --steps 1 --warmup 3 --no-cpu --no-sweep > /dev/null 2>&1
$NCU -k regex:ba_sweep_kernel -s 4 -c 2 -o gpurun_out/r02_ba_sweep python tools/ba_sweep_probe.py > /dev/null 2>&1
$NCU -k regex:"nlpf_tile_sweep|gather_grad" -s 6 -c 3 -o gpurun_out/r02_nlpf_sweep python tools/sweep_probe.py > /dev/null 2>&1
$NCU -k regex:"lm_" -s 40 -c 40 -o gpurun_out/r02_lm_dense python tools/lm_dense_probe.py 2 > /dev/null 2>&1
$NCU -k regex:"factor_bounds|list_bounds|cc_hook|cc_jump|solve_grid|solve_lm_block|solve_lm_ba_points|solve_strict" -c 12 -o gpurun_out/r02_misc python tools/misc_probe.py > /dev/null 2>&1
$NCU -k regex:solve_nlpf_resident -c 1 -o gpurun_out/r02_nlpf_resident python tools/cfg4_probe.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
