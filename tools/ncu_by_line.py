#!/usr/bin/env python
"""Correlates an `ncu --page source --csv` SASS listing (per-instruction 'Instructions Executed') with the source lines
of the kernel, using `nvdisasm -g` line annotations of the same cubin (instruction order is identical).
usage: ncu_by_line.py <source.csv> <nvdisasm -g -c output> <mangled kernel name> <source file>"""
import csv
import collections
import re
import sys

src_csv, dis, kname, cu = sys.argv[1:5]
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows[:6]) if "Source" in r)
hdr, data = rows[hi], rows[hi + 1:]
ia, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kname + ":"))
cur, inl, seq = None, None, []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section") or l.startswith(".section"):
        if seq:
            break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if m.group(1).endswith(cu):
            cur = int(m.group(2))
        elif m.group(3) and m.group(3).endswith(cu):
            cur = int(m.group(4))
        # else: a helper header (common.cuh): keep the last line of the kernel's own file (nearest-preceding attribution)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(1).strip()))
print("sass in ncu:", len(data), "sass in nvdisasm:", len(seq))
tot = sum(int(r[ia]) for r in data)
by = collections.Counter()
for (ln, txt), r in zip(seq, data):
    by[ln] += int(r[ia])
src = open("/root/repo/crixus_b200/csrc/" + cu).read().split("\n")
print("total warp instructions", tot)
if len(sys.argv) > 5:      # line ranges "a-b,c-d,..." -> share per range
    for rg in sys.argv[5].split(","):
        lo, hi_ = map(int, rg.split("-"))
        v = sum(c for l, c in by.items() if l is not None and lo <= l <= hi_)
        print("lines %4d-%4d %5.1f%%  %s" % (lo, hi_, 100.0 * v / tot, src[lo - 1].strip()[:90]))
for ln, v in sorted(by.items(), key=lambda x: -x[1])[:int(sys.argv[6]) if len(sys.argv) > 6 else 60]:
    print("%6d %5.1f%%  %s" % (ln, 100.0 * v / tot, src[ln - 1].strip()[:110] if ln and ln > 0 else "?"))
