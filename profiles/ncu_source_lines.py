"""Per-source-line hot spots from an ncu report (needs -lineinfo and --import-source on).
usage: python profiles/ncu_source_lines.py file.ncu-rep [top_n]
Prints, per CUDA source line: warp instructions executed, share of the kernel, average active threads, stall samples."""
import csv
import subprocess
import sys

src = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu', '-i', src, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows, fname, hdr = [], None, None
for r in csv.reader(txt.splitlines()):
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
    elif hdr and r[0].isdigit() and len(r) > 10:
        g = lambda name: r[hdr.index(name)]
        try:
            rows.append((fname, int(r[0]), r[1].strip()[:90], int(g('Instructions Executed')), int(g('Thread Instructions Executed')),
                         int(g('# Samples'))))
        except ValueError:
            pass
tot_i = sum(r[3] for r in rows) or 1
tot_s = sum(r[5] for r in rows) or 1
print(f'total warp instructions {tot_i:,}  thread instructions {sum(r[4] for r in rows):,}  avg active {sum(r[4] for r in rows)/tot_i:.2f}  samples {tot_s:,}')
for r in sorted(rows, key=lambda r: -r[5])[:top]:
    print(f'{r[0]:14s}:{r[1]:<4d} inst {100*r[3]/tot_i:5.1f}%  samples {100*r[5]/tot_s:5.1f}%  active {r[4]/max(r[3],1):5.1f}  | {r[2]}')
