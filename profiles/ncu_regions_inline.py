"""Like ncu_regions.py, but attributes every SASS instruction through its INLINE CHAIN (nvdisasm -gi on the cubin of the same build):
an instruction of an inlined helper (rt_math.cuh, the CUDA intrinsics headers) counts for the region of the innermost caller line that
lies in a region — so shading, the exact test and the pre-filter get their arithmetic back.
usage: python profiles/ncu_regions_inline.py file.ncu-rep regions.json path/to/librt_b200.so 'mangled kernel name'"""
import csv, json, os, re, subprocess, sys, tempfile

rep, regions, lib, kern = sys.argv[1], json.load(open(sys.argv[2])), sys.argv[3], sys.argv[4]
GENERIC = ('rt_math.cuh', 'sm_', 'device_', 'cuda_', 'math_', 'vector_', 'crt/')

def region_of(chain):
    for f, ln in chain:                                   # innermost first
        if f.startswith(GENERIC):
            continue
        for name, spans in regions.items():
            if any(f == s[0] and s[1] <= ln <= s[2] for s in spans):
                return name
    return 'other'

# ---- per-instruction inline chains from the cubin ----
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
chains, ops = [], []
for cub in sorted(os.listdir(tmp)):
    if 'sm_100' not in cub:
        continue
    txt = subprocess.run(['nvdisasm', '-gi', os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    key = f'.text.{kern}:'
    if key not in txt:
        continue
    body = txt.split(key, 1)[1]
    cur = []
    fresh = True
    for line in body.splitlines():
        if line.startswith('\t.section') or (line.startswith('.text.') and line.endswith(':')):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            if fresh:
                cur, fresh = [], False
            cur.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
        if m:
            chains.append(list(cur)); ops.append(m.group(2).strip()); fresh = True
    break
assert chains, 'kernel not found in the library'

# ---- per-instruction counters from the report ----
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows, hdr = [], None
for r in csv.reader(txt.splitlines()):
    if not r: continue
    if r[0] == 'Address': hdr = r
    elif hdr and r[0].startswith('0x') and len(r) > 6:
        g = lambda name: r[hdr.index(name)]
        rows.append((r[1].strip(), int(g('Instructions Executed')), int(g('Thread Instructions Executed')), int(g('# Samples'))))
assert len(rows) == len(chains), (len(rows), len(chains), 'the report and the library are different builds')
bad = sum(1 for (a, *_), b in zip(rows, ops) if a.split()[0].lstrip('@!P0123456789 ') [:4] != b.split()[0].lstrip('@!P0123456789 ')[:4])
acc = {k: [0, 0, 0] for k in list(regions) + ['other']}
for (op, wi, ti, sm), ch in zip(rows, chains):
    a = acc[region_of(ch)]; a[0] += wi; a[1] += ti; a[2] += sm
tw = sum(a[0] for a in acc.values()) or 1; ts = sum(a[2] for a in acc.values()) or 1
print(f"{len(rows)} SASS instructions, {bad} opcode mismatches between report and library")
print(f"{'region':34s} warp-inst %  samples %  active lanes")
for k, a in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:34s} {100*a[0]/tw:10.1f} {100*a[2]/ts:10.1f} {a[1]/max(a[0],1):10.1f}")
