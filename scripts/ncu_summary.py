"""Key metrics of an ncu report as a short text block (for profiles/).  python scripts/ncu_summary.py rep [title]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines())); hdr, units, vals = rows[0], rows[1], rows[2]
pat = re.compile(r'^(gpu__time_duration.sum|dram__bytes_(read|write).sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|'
                 r'sm__warps_active.avg.pct_of_peak_sustained_active|launch__(registers_per_thread|occupancy_limit_\w+|block_size|'
                 r'grid_size|shared_mem_per_block_dynamic|waves_per_multiprocessor)|smsp__inst_executed.sum|'
                 r'smsp__issue_active.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|'
                 r'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|lts__t_sectors_op_write.sum|lts__t_sectors_op_read.sum|'
                 r'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|'
                 r'smsp__average_warps_issue_stalled_\w+_per_issue_active.ratio|smsp__warp_issue_stalled_(barrier|long_scoreboard|short_scoreboard|lg_throttle|mio_throttle|membar|wait|not_selected|no_instruction|math_pipe_throttle|sleeping|dispatch_stall|branch_resolving|drain|imc_miss|tex_throttle)_per_warp_active.pct)$')
name = [v for h, v in zip(hdr, vals) if h == 'Kernel Name']
print('#', ' '.join(sys.argv[2:]) or rep)
if name: print('kernel,', name[0])
for h, u, v in zip(hdr, units, vals):
    if pat.search(h): print(f'{h},{u},{v}')
