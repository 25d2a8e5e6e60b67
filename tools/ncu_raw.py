"""Print the handful of raw ncu metrics the roofline discussion needs.
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_raw.py"""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__cycles_active.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__waves_per_multiprocessor',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg',
        'sm__inst_executed_pipe_tensor.sum', 'launch__cluster_size', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed']
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
for r in rows[2:]:
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w:
                print('  %-72s %s %s' % (h, r[i], rows[1][i]))
    print()
