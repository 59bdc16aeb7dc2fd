#!/usr/bin/env python
"""Aggregates an ncu report of one pass of a BASELINE workload into profiles/ncu_summary.json, the per-call figures bench.py
quotes beside the live timings (`roofline.traffic`, `issue_slot_frac`, `instr_per_vertex`, `executed_fma_frac`):

    python tools/ncu_to_json.py REPORT.ncu-rep config4 NVERT PASSES [profiles/ncu_summary.json]

PASSES = passes of the hot path in the report (tools/profile_pass.py runs two: set-up + one step); figures are per pass.

Kernels are grouped by the C-ABI call they belong to: "k_vert_volume" = everything cx_calc_vert_volume launches (k_vv_adj_*,
k_vv_fans, k_vv_rows<AX>, k_vv_lines<AX>, k_vert_volume), "fill_resolve" = the directed tests of the geometry fill with their
preparation (k_seg_*, k_cell*, k_block_list, k_resolve_*), "fill_flood" = the propagation rounds.  Issue-slot
and pipe fractions are duration-weighted means over the group's launches."""
import csv
import io
import json
import os
import subprocess
import sys

GROUPS = {"k_vert_volume": ("k_vv_", "k_vert_volume"),
          "fill_resolve": ("k_resolve_", "k_seg_prep", "k_seg_cells", "k_cell3_flags", "k_cell_planes", "k_block_list"),
          "fill_flood": ("k_fill_tiles", "k_build_list")}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    rep, key, nvert, passes = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4])
    out = sys.argv[5] if len(sys.argv) > 5 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_summary.json")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, scale_units=True):
        if name not in col:
            return 0.0
        v = num(r[col[name]])
        u = units[col[name]]
        if scale_units:
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3, "msecond": 1.0}.get(u, 1.0)
        return v
    res = {}
    for g, prefixes in GROUPS.items():
        sel = [r for r in rows[2:] if any(p in r[col["Kernel Name"]] for p in prefixes)]
        if not sel:
            continue
        ms = sum(get(r, "gpu__time_duration.sum") for r in sel) / passes
        inst = sum(get(r, "smsp__inst_executed.sum") for r in sel) / passes
        wsum = lambda name: sum(get(r, name) * get(r, "gpu__time_duration.sum") for r in sel) / max(ms * passes, 1e-30)  # noqa: E731
        res[g] = {
            "launches": len(sel) / passes, "duration_ms_under_ncu": ms,
            "dram_bytes": sum(get(r, "dram__bytes_read.sum") + get(r, "dram__bytes_write.sum") for r in sel) / passes,
            "warp_instr": inst, "warp_instr_per_vertex": inst / nvert if g == "k_vert_volume" else None,
            "issue_slot_frac": wsum("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
            "fma_pipe_frac": wsum("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") / 100.0,
            "alu_pipe_frac": wsum("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") / 100.0,
            "source": os.path.basename(rep),
        }
    data = {}
    if os.path.exists(out):
        data = json.load(open(out))
    data[key] = res
    json.dump(data, open(out, "w"), indent=1, sort_keys=True)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
