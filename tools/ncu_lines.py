#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS export per CUDA source line, using nvdisasm -g line info of the same cubin.
usage: ncu_lines.py <sass_csv> <nvdisasm -g -c output> <kernel mangled-name substring> [top]"""
import csv, re, sys, collections
sass_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# offset -> line from nvdisasm
off2line = {}
cur = None; inside = False
for l in open(dis):
    if l.startswith('\t.section') or l.startswith('//----'):
        inside = (kname in l) if '.text.' in l else inside
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m and inside:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
    m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(\S.*?);', l)
    if m and inside:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; data = rows[2:]
si = hdr.index('# Samples'); ai = hdr.index('Address'); ei = hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = int(data[0][ai], 16)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r in data:
    off = int(r[ai], 16) - base
    ln = off2line.get(off)
    n = int(r[si]); tot += n
    a = agg[ln]; a[0] += n; a[1] += int(r[ei])
    for c in stall_cols:
        a[2][hdr[c]] += int(r[c])
print('total samples', tot)
for ln, (n, ex, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/tot:5.1f}%  {n:6d}  exec {ex:9d}  {ln}  {[(k.replace('stall_',''),v) for k,v in st.most_common(2)]}")

# ---- role attribution: helper (inlined) lines inherit the role of the surrounding code in address order
if len(sys.argv) > 5:
    ranges = eval(sys.argv[5])  # {'producer': (a,b), ...} line ranges in the main file
    role = None
    racc = collections.defaultdict(lambda: [0, collections.Counter(), collections.Counter()])
    for r in data:
        off = int(r[ai], 16) - base
        ln = off2line.get(off)
        if ln and ln[0] == 'mlp_tc.cu':
            for k, (a, b) in ranges.items():
                if a <= ln[1] <= b:
                    role = k
        n = int(r[si])
        racc[role][0] += n
        racc[role][2][ln] += n
        for c in stall_cols:
            racc[role][1][hdr[c]] += int(r[c])
    for k, (n, st, lines) in racc.items():
        print(f"ROLE {k}: {100*n/tot:5.1f}%  {[(a.replace('stall_',''),b) for a,b in st.most_common(4)]}")
        print("     ", [(l, c) for l, c in lines.most_common(6)])
