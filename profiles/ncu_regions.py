"""Warp instructions, stall samples and active lanes of an ncu report, summed over named source-line regions.
usage: python profiles/ncu_regions.py file.ncu-rep regions.json   (regions: {"name": [["file", lo, hi], ...], ...}; first match wins)"""
import csv, json, subprocess, sys
src, regions = sys.argv[1], json.load(open(sys.argv[2]))
txt = subprocess.run(['ncu', '-i', src, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows, fname, hdr = [], None, None
for r in csv.reader(txt.splitlines()):
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]
    elif r[0] == 'Line No': hdr = r
    elif hdr and r[0].isdigit() and len(r) > 10:
        g = lambda name: r[hdr.index(name)]
        try: rows.append((fname, int(r[0]), int(g('Instructions Executed')), int(g('Thread Instructions Executed')), int(g('# Samples'))))
        except ValueError: pass
acc = {k: [0, 0, 0] for k in list(regions) + ['other']}
other = {}
for f, ln, wi, ti, sm in rows:
    hit = 'other'
    for name, spans in regions.items():
        if any(f == s[0] and s[1] <= ln <= s[2] for s in spans):
            hit = name; break
    a = acc[hit]; a[0] += wi; a[1] += ti; a[2] += sm
    if hit == 'other': other[(f, ln)] = other.get((f, ln), 0) + wi
tw = sum(a[0] for a in acc.values()) or 1; ts = sum(a[2] for a in acc.values()) or 1
print(f"{'region':28s} warp-inst %  samples %  active lanes")
for k, a in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} {100*a[0]/tw:10.1f} {100*a[2]/ts:10.1f} {a[1]/max(a[0],1):10.1f}")
print('largest unassigned lines:', sorted(other.items(), key=lambda kv: -kv[1])[:12])
