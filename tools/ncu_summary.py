"""Summarise an .ncu-rep (raw page) into the handful of metrics used in DESIGN.md / profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [('Kernel Name', 'kernel'), ('Grid Size', 'grid'), ('gpu__time_duration.sum', 'time'),
        ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
        ('lts__t_bytes.sum', 'l2_bytes'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct'),
        ('lts__t_sector_hit_rate.pct', 'l2_hit'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor_pct'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1_pct'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ_pct'),
        ('launch__registers_per_thread', 'regs'), ('sm__cycles_elapsed.max', 'cycles'),
        ('smsp__inst_executed.sum', 'inst'), ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem_wavefronts'),
        ('smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'stall_long_sb'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall_long_sb2'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_pct')]
for r in data:
    print('---')
    for h, nm in want:
        if h in col:
            v = r[col[h]]
            if nm == 'kernel':
                v = v[:60]
            print(f'  {nm:16s} {v} {units[col[h]]}')
