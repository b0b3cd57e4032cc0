"""Print the metrics we track from an .ncu-rep (raw page) - used to write profiles/*.md."""
import csv, subprocess, sys
rep = sys.argv[1]
rows = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
H = rows[0]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max', 'launch__shared_mem_per_block_dynamic']
for r in rows[2:]:
    for w in want:
        for i, h in enumerate(H):
            if h == w or (w.startswith('sm__pipe_tensor') and h.startswith('sm__pipe_tensor') and 'pct_of_peak_sustained_active' in h):
                print(f'{h:80s} {r[i]} {rows[1][i]}')
    st = {}
    for i, h in enumerate(H):
        if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
            try: st[h.replace('smsp__pcsamp_warps_issue_stalled_', '')] = float(r[i].replace(',', ''))
            except: pass
    tot = sum(st.values()) or 1
    print('stalls:', ', '.join(f'{k} {v/tot*100:.0f}%' for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    print()
