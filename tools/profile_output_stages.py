#!/usr/bin/env python
"""Times the stages either side of the hot path on BASELINE configs[1] (device resident, CUDA events on the context's
stream): special-boundary tagging (two wall-sized rectangles + 96 adversarial triangles), particle-record assembly, the VTU
body kernel + streamed write to /tmp.  Prints algorithmic bytes / time against the measured HBM peak (MEASURED_PEAKS.json).

    gpurun -- python tools/profile_output_stages.py gpurun_out/output_stages.json [levels]
    ncu --set full -k regex:'k_rec_|k_sb_|k_vtu' ... python tools/profile_output_stages.py /dev/null"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import cases  # noqa: E402
from crixus_b200 import api, workloads as wl  # noqa: E402
from crixus_b200.dist import ShardedHotPath  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for k in range(reps):
        fn()
        ev[k + 1].record()
    torch.cuda.synchronize()
    return min(ev[k].elapsed_time(ev[k + 1]) for k in range(reps))


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "output_stages.json")
    levels = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6539.9
    ctx = api.Context(0)
    w = wl.config2(levels)
    hp = ShardedHotPath(ctx, w, 0, 1, byte_stats=False)
    nv, nbe, dev = hp.nv, hp.nbe, ctx.device
    res = {"workload": w["name"], "triangles": nbe, "vertices": nv, "hbm_peak_gbs": hbm}
    # ---- special boundaries
    tris = np.asarray(w["tris"], dtype=np.float32)
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    y0, y1, z0, z1 = lo[1] + 0.2 * (hi[1] - lo[1]), lo[1] + 0.7 * (hi[1] - lo[1]), lo[2] + 0.1 * (hi[2] - lo[2]), lo[2] + 0.5 * (hi[2] - lo[2])
    walls = np.array([[[x, y0, z0], [x, y1, z0], [x, y1, z1]] for x in (lo[0], hi[0])] +
                     [[[x, y0, z0], [x, y1, z1], [x, y0, z1]] for x in (lo[0], hi[0])], dtype=np.float32)
    grids = [walls, cases.special_stress(tris, 3, n=96)]
    sbid = torch.zeros(nv + nbe, dtype=torch.int32, device=dev)
    dgrids = []
    for g in grids:
        sv, se = api.weld_host(g.reshape(-1, 3), np.float32(w["dr"]))
        sp = np.zeros((sv.shape[0], 4), dtype=np.float32); sp[:, :3] = sv
        e4 = np.zeros((se.shape[0], 4), dtype=np.int32); e4[:, :3] = se
        dgrids.append((torch.from_numpy(sp).to(dev), torch.from_numpy(e4).to(dev), e4.shape[0]))

    def tag():
        sbid.zero_()
        for k, (sp, e4, n) in enumerate(dgrids):
            ctx.identify_special_boundary(hp.pf, hp.pos, hp.ep, hp.cell_idx, nv, nbe, sp, e4, n, sbid, k + 1)
    ms = timed(tag)
    tagged = int((sbid[nv:] > 0).sum())
    res["special_boundary"] = {"ms": ms, "grid_triangles": [g[2] for g in dgrids], "segments_tagged": tagged,
                               "reference_pair_tests": int(nbe) * sum(g[2] for g in dgrids)}
    # ---- record assembly (through the cx_job view of the device arrays)
    job = api.CxJob()
    job.nvert, job.nbe, job.nfluid = nv, nbe, int(hp.nfluid)
    job.params_fill = hp.pf
    for j in range(3):
        job.dimg[j] = int(hp.dimg[j])
    job.d_pos, job.d_norm, job.d_ep, job.d_surf = hp.pos.data_ptr(), hp.norm.data_ptr(), hp.ep.data_ptr(), hp.surf.data_ptr()
    job.d_vol, job.d_fpos, job.d_sbid = hp.vol.data_ptr(), hp.fpos.data_ptr(), sbid.data_ptr()
    rec = ctx.assemble_records_device(job)
    n = int(rec.shape[0])
    del rec                                               # (two live 0.74 GB buffers would put a cudaMalloc into the timing)
    t0 = time.perf_counter()
    for _ in range(3):
        rec = ctx.assemble_records_device(job)
        del rec
    asm_ms = (time.perf_counter() - t0) / 3 * 1e3          # host clock: the call synchronises to read the counts back
    nwords = int(hp.fpos.numel())
    alg = 64 * n + 4 * nwords + 16 * (nv + nbe) + 4 * nv + (16 + 16 + 4 + 4 + 12) * nbe + 4 * nv
    res["assemble_records"] = {"ms": asm_ms, "particles": n, "algorithmic_bytes": alg, "gbs": alg / asm_ms / 1e6,
                               "hbm_frac": alg / asm_ms / 1e6 / hbm, "note": "ids present: flag-scan per id included"}
    job.d_sbid = None
    t0 = time.perf_counter()
    for _ in range(3):
        rec = ctx.assemble_records_device(job)
        del rec
    asm0_ms = (time.perf_counter() - t0) / 3 * 1e3
    rec = ctx.assemble_records_device(job)
    alg0 = 64 * n + 4 * nwords + 16 * (nv + nbe) + 4 * nv + (16 + 16 + 4 + 12) * nbe
    res["assemble_records_no_ids"] = {"ms": asm0_ms, "algorithmic_bytes": alg0, "gbs": alg0 / asm0_ms / 1e6, "hbm_frac": alg0 / asm0_ms / 1e6 / hbm}
    for fmt, per in (("vtu", 84), ("h5sph", 64)):
        fn = "/tmp/cx_profile_out." + fmt
        t0 = time.perf_counter()
        ctx.write_records_device(rec, fn, fmt)
        dt = time.perf_counter() - t0
        res["write_" + fmt] = {"s": dt, "bytes": os.path.getsize(fn), "gb_per_s": os.path.getsize(fn) / dt / 1e9}
        os.remove(fn)
    print(json.dumps(res, indent=1))
    if out != "/dev/null":
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
