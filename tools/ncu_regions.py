#!/usr/bin/env python
"""Per source REGION of one kernel: SASS size, warp instructions executed and stall samples (all / no_instruction), from an
`ncu --page source --csv --print-source sass` export joined with `nvdisasm -g -c` line info of the same cubin.  Instructions of
inlined helpers (regions given as name=None) are attributed to the last named region before them in address order.

    ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv      (one kernel per export: use --kernel-name / -k)
    python tools/ncu_regions.py x.csv vvg.txt '<mangled kernel>' vert_volume.cu 'A.setup:682-740,A.gather:741-765,...' [launch#]
Regions not listed are 'helper' (ambiguous)."""
import csv
import re
import sys


def main():
    src_csv, dis, kname, cu, spec = sys.argv[1:6]
    which = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    regions = []
    for part in spec.split(","):
        name, rng = part.split(":")
        a, b = rng.split("-")
        regions.append((int(a), int(b), name))
    # split the csv into per-launch tables
    rows = list(csv.reader(open(src_csv)))
    tables, cur = [], None
    for r in rows:
        if len(r) > 3 and "Source" in r and "Address" in r:
            cur = {"hdr": r, "data": []}
            tables.append(cur)
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    t = tables[which]
    hdr, data = t["hdr"], t["data"]
    ia = hdr.index("Instructions Executed")
    isamp = hdr.index("# Samples")
    inoi = hdr.index("stall_no_inst")
    lines = open(dis).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kname + ":"))
    ctx, seq = "?", []
    for l in lines[start + 1:]:
        if l.startswith(".text.") and seq:
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            if m.group(1).endswith(cu):
                n = int(m.group(2))
                for a, b, name in regions:
                    if a <= n <= b:
                        ctx = name
                        break
            continue
        if re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l):
            seq.append(ctx)
    assert len(seq) == len(data), (len(seq), len(data))
    agg = {}
    for c, r in zip(seq, data):
        e = agg.setdefault(c, [0, 0, 0, 0])
        e[0] += 1
        e[1] += int(r[ia] or 0)
        e[2] += int(r[isamp] or 0)
        e[3] += int(r[inoi] or 0)
    tot = [sum(v[i] for v in agg.values()) for i in range(4)]
    print("%-14s %7s %7s %9s %9s %9s %12s" % ("region", "SASS", "KB", "instr %", "samples %", "no_inst %", "no_inst/samp"))
    for c, v in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print("%-14s %7d %7.1f %9.1f %9.1f %9.1f %12.2f" % (c, v[0], v[0] * 16 / 1024, 100 * v[1] / tot[1], 100 * v[2] / tot[2],
                                                            100 * v[3] / max(tot[3], 1), v[3] / max(v[2], 1)))
    print("%-14s %7d %7.1f %9d %9d %9d %12.2f" % ("total", tot[0], tot[0] * 16 / 1024, tot[1], tot[2], tot[3], tot[3] / max(tot[2], 1)))


if __name__ == "__main__":
    main()
