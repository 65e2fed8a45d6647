"""Annotated SASS of a kernel with executed warp-instructions per unit (e.g. per trial move) from an ncu source-page csv.
usage: ncu_hot_sass.py <src.csv> <nvdisasm -g -c output> <mangled kernel> <units per launch> [min per unit]   (development aid)"""
import csv, re, sys
src_csv, dis_txt, k, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 0.05
lines = open(dis_txt, errors='replace').read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + k + ':'))
seq = []; cur = ('?', 0)
for l in lines[start + 1:]:
    if l.startswith('//-------') or l.startswith('\t.section'): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l): seq.append((cur, l.strip()))
    elif re.match(r"\.L_x_\d+:", l.strip()): seq.append((cur, 'LABEL ' + l.strip()))
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]; ci = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed"); si = hdr.index("# Samples")
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
i = 0
for (cur, ins) in seq:
    if ins.startswith('LABEL'):
        print(ins); continue
    r = data[i]; i += 1
    n = int(float(r[ci] or 0)); t = int(float(r[ti] or 0)); s = int(float(r[si] or 0))
    if n / units >= thr:
        print(f"{n/units:6.2f} {t/max(n,1):5.1f} {s:6d} {cur[0]}:{cur[1]:<4d} {ins[:110]}")
