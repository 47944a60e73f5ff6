import csv, sys, subprocess, collections
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
h=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','launch__grid_size','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']
for w in want:
    if w in h: print(w, rows[1][h.index(w)], rows[2][h.index(w)])
out=[]
for i,c in enumerate(h):
    if 'pcsamp_warps_issue_stalled' in c and 'not_issued' not in c:
        try: out.append((float(rows[2][i].replace(',','')), c.replace('smsp__pcsamp_warps_issue_stalled_','')))
        except: pass
tot=sum(v for v,_ in out)
print(' '.join(f"{c}={100*v/tot:.0f}%" for v,c in sorted(out,reverse=True)[:8]))
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hi = next(i for i,r in enumerate(rows) if '# Samples' in r)
h=rows[hi]; ie=h.index('Instructions Executed'); si=h.index('# Samples')
ops=collections.Counter(); n=0; seen=set()
for r in rows[hi+1:]:
    try: v=int(r[ie])
    except: continue
    if r[0] in seen: continue
    seen.add(r[0])
    t=r[1].strip().split()
    op=t[1] if t[0].startswith('@') else t[0]
    ops[op.split('.')[0]]+=v; n+=v
print('total inst',n, ' '.join(f"{op}={100*v/n:.1f}%" for op,v in ops.most_common(14)))
