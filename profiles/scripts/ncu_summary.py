"""Summarise an .ncu-rep (read here, no GPU): python profiles/scripts/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max', 'launch__shared_mem_per_block_dynamic', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
for r in rows[2:]:
    d = dict(zip(h, r))
    print(d.get('Kernel Name'), d.get('Grid Size'), d.get('Block Size'))
    for k in h:
        if k in want:
            print(f"    {k} {d[k]} {rows[1][h.index(k)]}")
    st = {k.split('issue_stalled_')[1].split('_per_')[0]: float(d[k]) for k in h if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('per_issue_active.ratio')}
    for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
        print(f"    stall {k} {v:.2f}")
