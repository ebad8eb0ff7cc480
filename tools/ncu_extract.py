import csv, sys
rep, out = sys.argv[1], sys.argv[2]
import subprocess
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name','gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size',
 'dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__average_warp_latency_per_inst_issued.ratio']
with open(out, 'w') as f:
    for i, h in enumerate(hdr):
        if h in want or ('issue_stalled' in h and 'per_issue_active' in h) or 'pipe_fp64' in h or 'pipe_fmaheavy' in h:
            f.write(f"{h},{units[i]},{vals[i]}\n")
