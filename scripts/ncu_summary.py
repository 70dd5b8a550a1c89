"""Summarises an .ncu-rep (raw page) into a small CSV + prints key metrics.  Usage: ncu_summary.py rep out.csv"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
keep = ['ID', 'Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg', 'TPC.TriageCompute.sm__cycles_active.avg', 'gpc__cycles_elapsed.max', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_barrier',
        'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_selected',
        'smsp__pcsamp_warps_issue_stalled_not_selected', 'smsp__pcsamp_warps_issue_stalled_mio_throttle', 'smsp__pcsamp_warps_issue_stalled_dispatch_stall',
        'smsp__pcsamp_warps_issue_stalled_lg_throttle', 'smsp__pcsamp_warps_issue_stalled_no_instructions']
keep = [k for k in keep if k in idx]
with open(out, 'w', newline='') as f:
    w = csv.writer(f); w.writerow(keep); w.writerow([units[idx[k]] for k in keep])
    for d in data: w.writerow([d[idx[k]] for k in keep])
for k in keep[2:]:
    print(f"{k[:70]:70s} {units[idx[k]]:10s}", [d[idx[k]][:9] for d in data[:8]])
