"""Per-source-line summary of an ncu report: python tools/ncu_lines.py report.ncu-rep [top]  (needs -lineinfo + --import-source on)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname, hdr, agg = '', None, []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif len(r) > 5 and r[0] == 'Line No':
        hdr = r
        i_s, i_e = hdr.index('# Samples'), hdr.index('Instructions Executed')
    elif hdr and len(r) > 10 and r[2] == '-':      # a source line (its SASS rows follow, with an address in column 2)
        try:
            agg.append((fname, int(r[0]), r[1], int(r[i_s]), int(r[i_e])))
        except ValueError:
            pass
ts, te = sum(a[3] for a in agg) or 1, sum(a[4] for a in agg) or 1
print(f'samples {ts}  warp instructions {te}')
for f, l, src, s, e in sorted(agg, key=lambda a: -a[4])[:top]:
    print(f'{f[:20]:20s} {l:5d}  exec {e / te * 100:5.2f}%  samples {s / ts * 100:5.2f}%  {src.strip()[:105]}')
