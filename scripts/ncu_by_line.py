"""Aggregate an ncu SASS-page csv (per-instruction executed counts / stall samples) by CUDA source
line, using nvdisasm -g line markers of the same cubin.  Development aid.
usage: ncu_by_line.py <src.csv from `ncu --page source --csv`> <nvdisasm -g -c output> <mangled kernel name> [top]"""
import csv, re, sys
from collections import defaultdict

src_csv, dis_txt, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
# ---- nvdisasm: sequence of (file,line) per instruction inside the kernel's .text section
lines = open(dis_txt, errors="replace").read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kname + ":"))
seq = []
cur = ("?", 0)
inl = ""
for l in lines[start + 1:]:
    if l.startswith("//-------") or l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l):
        seq.append(cur)
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
print("sass instrs: ncu", len(data), "nvdisasm", len(seq))
n = min(len(data), len(seq))
agg = defaultdict(lambda: [0, 0, 0, 0])
tot = 0; tots = 0
for k in range(n):
    ie = int(float(data[k][ci] or 0)); sm = int(float(data[k][cs] or 0)); te = int(float(data[k][ct] or 0))
    a = agg[seq[k]]
    a[0] += ie; a[1] += sm; a[2] += te; a[3] += 1
    tot += ie; tots += sm
print("total warp-instr", tot, "samples", tots)
srcs = {}
def srcline(f, ln):
    import os
    for d in ("/root/repo/mc_water_ls_mw_b200/csrc/",):
        p = d + f
        if os.path.exists(p):
            if p not in srcs: srcs[p] = open(p).read().splitlines()
            if 0 < ln <= len(srcs[p]): return srcs[p][ln - 1].strip()[:90]
    return ""
items = sorted(agg.items(), key=lambda kv: -kv[1][0])
print(f"{'inst%':>6} {'smp%':>6} {'lanes':>5} {'sass':>5}  where")
for (f, ln), (ie, sm, te, ns) in items[:top]:
    print(f"{ie/tot*100:6.2f} {sm/max(tots,1)*100:6.2f} {te/max(ie,1):5.1f} {ns:5d}  {f}:{ln}  {srcline(f, ln)}")
# per-file-function coarse buckets

# ---- coarse buckets by (file, line range): edit RANGES to taste
import os
RANGES = os.environ.get("NCU_RANGES")
if RANGES:
    buckets = []
    for item in RANGES.split(";"):
        name, f, a, b = item.split(":")
        buckets.append((name, f, int(a), int(b)))
    sums = defaultdict(lambda: [0, 0])
    other = [0, 0]
    for (f, ln), (ie, sm, te, ns) in agg.items():
        for name, bf, a, b in buckets:
            if f == bf and a <= ln <= b:
                sums[name][0] += ie; sums[name][1] += sm
                break
        else:
            other[0] += ie; other[1] += sm
            sums["other:" + f][0] += ie; sums["other:" + f][1] += sm
    print("\nbuckets:")
    for name, (ie, sm) in sorted(sums.items(), key=lambda kv: -kv[1][0]):
        print(f"{ie/tot*100:6.2f}% inst {sm/max(tots,1)*100:6.2f}% smp  {name}")
