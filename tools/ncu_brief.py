"""Dev tool: the handful of ncu raw metrics the kernel notes quote.  usage: python tools/ncu_brief.py report.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts.sum', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_lg.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed', 'l1tex__f_wavefronts.avg.pct_of_peak_sustained_elapsed' ]
for i, x in enumerate(h):
    if x in want or ('issue_stalled' in x and 'per_issue_active' in x and float(v[i] or 0) > 0.15) or x.startswith('smsp__inst_executed_pipe_') and x.endswith('.sum'):
        print(x, u[i], v[i])
