#!/bin/bash
# Round-2 evidence capture (run under gpurun, ONE GPU): launch list of the bench step + ncu --set full of every kernel family.
# The reports (25-80 MB each) are summarised ON the box (tools/ncu_raw.py, tools/ncu_stalls.py) and deleted: gpurun brings
# back at most 64 MiB.  Summaries land in gpurun_out/r02_ncu_*.txt and are copied to profiles/ by hand.
NCU="ncu --set full --clock-control none --import-source on"
summarise() {  # <name>
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw.py > gpurun_out/r02_ncu_$1_raw.txt 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_stalls.py > gpurun_out/r02_ncu_$1_stalls.txt 2>&1
  rm -f gpurun_out/$1.ncu-rep
}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep > /dev/null 2>&1
$NCU -k regex:"solve_ba_(cameras|points)" -s 6 -c 2 -o gpurun_out/solve_ba python bench.py --steps 1 --warmup 3 --no-cpu --no-sweep > /dev/null 2>&1; summarise solve_ba
$NCU -k regex:ba_sweep_kernel -s 4 -c 2 -o gpurun_out/ba_sweep python tools/ba_sweep_probe.py > /dev/null 2>&1; summarise ba_sweep
$NCU -k regex:"nlpf_tile_sweep|gather_grad" -s 6 -c 3 -o gpurun_out/nlpf_sweep python tools/sweep_probe.py > /dev/null 2>&1; summarise nlpf_sweep
$NCU -k regex:"lm_" -s 40 -c 40 -o gpurun_out/lm_dense python tools/lm_dense_probe.py 2 > /dev/null 2>&1; summarise lm_dense
$NCU -k regex:"factor_bounds|list_bounds|cc_hook|cc_jump|solve_grid|solve_lm_block|solve_lm_ba_points|solve_strict" -c 12 -o gpurun_out/misc python tools/misc_probe.py > /dev/null 2>&1; summarise misc
$NCU -k regex:solve_nlpf_resident -c 1 -o gpurun_out/nlpf_resident python tools/cfg4_probe.py > /dev/null 2>&1; summarise nlpf_resident
ls -la gpurun_out/
