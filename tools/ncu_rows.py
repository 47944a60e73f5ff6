"""Key metrics + stall reasons of EVERY launch in an ncu report (tools/ncu_summary.py reads the first one only):
python tools/ncu_rows.py report.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    print('---')
    for w in want:
        if w in h:
            print(' ', w, r[h.index(w)][:90], rows[1][h.index(w)])
    out = []
    for i, c in enumerate(h):
        if 'pcsamp_warps_issue_stalled' in c and 'not_issued' not in c:
            try:
                out.append((float(r[i].replace(',', '')), c.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError:
                pass
    tot = sum(v for v, _ in out) or 1.0
    print('  ' + ' '.join(f"{c}={100 * v / tot:.0f}%" for v, c in sorted(out, reverse=True)[:6]))
