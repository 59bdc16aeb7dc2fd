"""Code layout of one kernel from `nvdisasm -g -c` (line info): which source region each stretch of SASS belongs to, in
address order, and the total per region.  Helper functions that are inlined in many places (mono_search, mono_mask, plane_kill,
intrinsics headers) are attributed to the last unambiguous region before them.

    cuobjdump -xelf all crixus_b200/libcrixus_b200.so && nvdisasm -g -c vert_volume.sm_100a.cubin > vvg.txt
    python tools/sass_layout.py vvg.txt '_Z13k_vert_volumeILi16ELb0ELb0ELb0EEv6VVArgs' crixus_b200/csrc/vert_volume.cu
"""
import re
import sys

REGIONS_VV = [  # (first line, last line, name) of vert_volume.cu; helper ranges are None = ambiguous
    (67, 203, None), (205, 258, "exact_tri"), (260, 322, None), (324, 376, "rows_dead"), (378, 405, "group_weight"),
    (407, 424, "lw.prologue"), (425, 472, "lw.loop1(V/E)"), (473, 541, "lw.loop2(W)"), (542, 583, "lw.cubes"),
    (585, 648, "A.setup"), (649, 672, "A.gather"), (673, 703, "A.chain"), (704, 731, "A.edges"), (732, 778, "A.planes"),
    (779, 813, "B.driver"), (814, 821, "B.final"),
]


def main():
    path, fn, src = sys.argv[1], sys.argv[2], sys.argv[3]
    regions = REGIONS_VV
    inside = False
    ctx = "?"
    line_re = re.compile(r'//## File "(.*)", line (\d+)')
    ins_re = re.compile(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?);")
    segs = []   # [ctx, first addr, last addr, n]
    amb = {}    # helper instructions per attributed ctx
    cur_amb = False
    for ln in open(path):
        if ln.startswith(".text."):
            inside = ln.strip() == ".text." + fn + ":"
            continue
        if not inside:
            continue
        m = line_re.search(ln)
        if m:
            if m.group(1).endswith(src.split("/")[-1]):
                n = int(m.group(2))
                cur_amb = True
                for a, b, name in regions:
                    if a <= n <= b:
                        if name is not None:
                            ctx, cur_amb = name, False
                        break
            else:
                cur_amb = True
            continue
        m = ins_re.match(ln)
        if m:
            addr = int(m.group(1), 16)
            if segs and segs[-1][0] == ctx:
                segs[-1][2] = addr
                segs[-1][3] += 1
            else:
                segs.append([ctx, addr, addr, 1])
            if cur_amb:
                amb[ctx] = amb.get(ctx, 0) + 1
    tot = {}
    for c, a, b, n in segs:
        tot[c] = tot.get(c, 0) + n
    # merge tiny interleavings for the layout view
    print("# layout (address order; stretches < 24 instructions folded into their predecessor)")
    merged = []
    for c, a, b, n in segs:
        if merged and (n < 24 or merged[-1][0] == c):
            merged[-1][2] = b
            merged[-1][3] += n
        else:
            merged.append([c, a, b, n])
    for c, a, b, n in merged:
        print(f"  {a:#08x}-{b:#08x} {n:5d} instr {n*16/1024:6.1f} KB  {c}")
    print("# totals (instructions, KB, of which inlined helpers)")
    for c, n in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"  {c:16s} {n:5d} {n*16/1024:6.1f} KB  helpers {amb.get(c, 0)}")
    print(f"  {'all':16s} {sum(tot.values()):5d} {sum(tot.values())*16/1024:6.1f} KB")


if __name__ == "__main__":
    main()
