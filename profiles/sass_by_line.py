"""Joins an `ncu --page source --print-source sass --csv` export of one kernel with `nvdisasm -g` line info and
sums warp-level and thread-level instruction counts per source line / per named region of rtfs_core.cuh.

usage: sass_by_line.py <src_sass.csv> <first row of the kernel block (1-based line of its "Kernel Name" row)> <all.sass> <mangled kernel name>
"""
import csv
import re
import sys
from collections import defaultdict

csv_path, start, sass_path, kernel = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]

# ---- line info from nvdisasm -g ----
lines = open(sass_path, errors="replace").read().split("\n")
begin = next(i for i, l in enumerate(lines) if l.startswith(".text." + kernel + ":"))
loc = []
cur = ("?", 0)
for l in lines[begin + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        loc.append((cur, l.split("*/", 1)[1].strip()))

# ---- counters from ncu ----
rows = list(csv.reader(open(csv_path)))
hdr = rows[start]  # row after "Kernel Name"
ia, it, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
body = []
for r in rows[start + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) > it:
        body.append(r)
assert len(body) == len(loc), (len(body), len(loc))

REGIONS = []  # optional: file with "first last label" lines for rtfs_core.cuh, argv[6]
if len(sys.argv) > 6:
    for l in open(sys.argv[6]):
        if l.strip() and not l.startswith("#"):
            a, b, label = l.split(None, 2)
            REGIONS.append((int(a), int(b), label.strip()))
per_line = defaultdict(lambda: [0, 0, 0])
for (fl, sass), r in zip(loc, body):
    a, t, s = int(r[ia]), int(r[it]), int(r[ist] or 0)
    e = per_line[fl]
    e[0] += a
    e[1] += t
    e[2] += s
tot_a = sum(v[0] for v in per_line.values())
tot_t = sum(v[1] for v in per_line.values())
tot_s = sum(v[2] for v in per_line.values())
print(f"total warp-instr {tot_a:.4g}  thread-instr {tot_t:.4g}  avg active {tot_t / tot_a:.2f}  samples {tot_s}")
print(f"{'file:line':34s} {'warp-instr %':>12s} {'avg lanes':>9s} {'samples %':>9s}")
for fl, v in sorted(per_line.items(), key=lambda x: -x[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 60]:
    print(f"{fl[0] + ':' + str(fl[1]):34s} {100 * v[0] / tot_a:12.2f} {v[1] / max(1, v[0]):9.2f} {100 * v[2] / max(1, tot_s):9.2f}")

if REGIONS:
    reg = defaultdict(lambda: [0, 0, 0])
    for fl, v in per_line.items():
        label = "other (" + fl[0] + ")"
        if fl[0] == "rtfs_core.cuh":
            for a, b, lab in REGIONS:
                if a <= fl[1] <= b:
                    label = lab
                    break
        e = reg[label]
        for k in range(3):
            e[k] += v[k]
    print()
    print(f"{'region':44s} {'warp-instr %':>12s} {'thread-instr %':>14s} {'avg lanes':>9s} {'samples %':>9s}")
    for lab, v in sorted(reg.items(), key=lambda x: -x[1][0]):
        print(f"{lab:44s} {100 * v[0] / tot_a:12.2f} {100 * v[1] / tot_t:14.2f} {v[1] / max(1, v[0]):9.2f} {100 * v[2] / max(1, tot_s):9.2f}")
