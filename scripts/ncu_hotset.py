"""Dynamic instruction footprint of a kernel from an ncu source-page csv (development aid).
usage: ncu_hotset.py <src.csv> <nvdisasm -g -c output> <mangled kernel> <units (e.g. moves) per launch>"""
import csv, re, sys
from collections import defaultdict
import numpy as np
src_csv, dis_txt, k, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
lines = open(dis_txt, errors='replace').read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + k + ":"))
seq = []; cur = ('?', 0)
for l in lines[start + 1:]:
    if l.startswith("//-------") or l.startswith("\t.section"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l): seq.append(cur)
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]; ci = hdr.index("Instructions Executed")
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
ex = np.array([int(float(r[ci] or 0)) for r in data])
print("static instrs", len(ex), " dynamic per unit %.0f" % (ex.sum() / units))
for thr in (1.0, 0.5, 0.2, 0.05, 0.01):
    n = (ex >= thr * units).sum()
    print("  executed >= %.2f x units: %5d instrs = %5.1f KB" % (thr, n, n * 16 / 1024))
b = defaultdict(lambda: [0, 0])
for kk in range(min(len(seq), len(ex))):
    if ex[kk] >= 0.2 * units:
        f, ln = seq[kk]; key = (f, ln // 10 * 10); b[key][0] += 1; b[key][1] += ex[kk]
print("hot (>=0.2/unit) static instrs by source region: count, dyn/unit")
for key, v in sorted(b.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 40]:
    print("  %4d %7.1f  %s:%d" % (v[0], v[1] / units, key[0], key[1]))
