#!/usr/bin/env python
"""Where does the z-slab fill spend its time?  Emulates N ranks on ONE GPU (one context + bitfield copy per rank, halo
planes exchanged by hand, like tests/test_gpu_properties.py) on the closed tank of BASELINE configs[4]'s family and times
every cx_fill_geometry_slab call: per outer iteration the slowest rank is the critical path of the real N-GPU run.

    gpurun -- python tools/slab_fill_profile.py 1000 8 [rounds_per_exchange]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402

from crixus_b200 import api  # noqa: E402
from crixus_b200.dist import ShardedHotPath, split_range  # noqa: E402
from run_large import build  # noqa: E402


def main():
    size, n = int(sys.argv[1]), int(sys.argv[2])
    rpe = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    v, t, nrm, w = build("tank", size)
    ctx = api.Context(0)
    corners = np.ascontiguousarray(v[t], dtype=np.float32).reshape(-1, 3)
    pos4, ep4, lo, hi = ctx.weld_device(torch.from_numpy(corners).to(ctx.device), w["dr"])
    norm4 = np.zeros((t.shape[0], 4), dtype=np.float32)
    norm4[:, :3] = nrm
    w.update(nvert=int(pos4.shape[0]), nbe=int(t.shape[0]), pos=pos4.cpu().numpy(), ep=ep4.cpu().numpy(), norm=norm4, bbmin=lo, bbmax=hi)
    hp = ShardedHotPath(ctx, w, 0, 1, byte_stats=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hp._fill()
    torch.cuda.synchronize()
    single_ms = (time.perf_counter() - t0) * 1e3
    f = [x for x in w["fills"] if x[0] == "geometry"][0]
    p = hp.pf.copy()
    p.dr_wall = np.float32(f[2])
    seed = api.seed_index(p, f[1])
    dimg = [int(x) for x in hp.dimg]
    P = dimg[0] * dimg[1]
    f0 = hp.fpos.clone()
    ctxs = [api.Context(0) for _ in range(n)]
    ks = [split_range(dimg[2], n, r) for r in range(n)]
    for c in ctxs:
        c.fill_set_max_rounds(rpe)

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return r, (time.perf_counter() - t0) * 1e3

    def run(form):
        """one emulated N-rank fill -> (critical path ms, per-call times, result check)"""
        fs = [torch.zeros_like(f0) for _ in range(n)]
        res = {}
        crit = 0.0
        if form == "dealt":
            views, ms = [], []
            for r in range(n):
                v_, t_ = timed(lambda: ctxs[r].fill_resolve_part(p, fs[r], hp.norm, hp.ep, hp.pos, hp.nbe, 0, hp.cell_idx, n, r))
                views.append(v_); ms.append(t_)
            tot = views[0].clone()
            for r in range(1, n):
                tot += views[r]
            for r in range(n):
                views[r].copy_(tot)
            res["resolve_part_ms"] = [round(x, 2) for x in ms]
            res["allreduce_bytes"] = int(tot.numel()) * 4
            crit += max(ms)
            del tot
        first, log, iters = True, [], 0
        while True:
            changed, unfinished, ms = 0, 0, []
            for r in range(n):
                fc = (2 if form == "dealt" else 1) if first else 0
                c_, t_ = timed(lambda: ctxs[r].fill_geometry_slab(p, fs[r], hp.norm, hp.ep, hp.pos, hp.nbe, 0, hp.cell_idx, seed, ks[r][0], ks[r][1], fc))
                changed += c_; ms.append(t_)
                unfinished += ctxs[r].fill_unfinished()
            first = False
            iters += 1
            crit += max(ms)
            log.append([round(x, 2) for x in ms])
            if changed == 0 and unfinished == 0:
                break
            for r in range(n - 1):
                for src, dst, k in ((r, r + 1, ks[r][1] - 1), (r + 1, r, ks[r + 1][0])):
                    a, b = (k * P) // 32, ((k + 1) * P + 31) // 32
                    fs[dst][a:b] |= fs[src][a:b]
        merged = torch.zeros_like(f0)
        for r in range(n):
            lo_b, hi_b = ks[r][0] * P, ks[r][1] * P
            part = fs[r].clone()
            w0, w1 = lo_b // 32, (hi_b + 31) // 32
            part[:w0] = 0
            part[w1:] = 0
            if lo_b % 32:
                part[w0] &= int(np.array((0xFFFFFFFF << (lo_b % 32)) & 0xFFFFFFFF, dtype=np.uint32).view(np.int32))
            if hi_b % 32:
                part[w1 - 1] &= int(np.array((1 << (hi_b % 32)) - 1, dtype=np.uint32).view(np.int32))
            merged |= part
        res.update(outer_iterations=iters, critical_path_ms=crit, first_call_ms=log[0], later_calls_max_ms=[max(x) for x in log[1:]],
                   equals_single_gpu=bool(torch.equal(merged, f0)))
        return res

    out = {"size": size, "lattice": dimg, "emulated_ranks": n, "rounds_per_exchange": rpe, "single_gpu_fill_ms": single_ms,
           "note": "every form is run twice per context; the second run (no cudaMalloc of the fill state) is reported"}
    for form in ("slab", "dealt"):
        run(form)
        out[form] = run(form)
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "slab_fill_%d_n%d_r%d.json" % (size, n, rpe)), "w"), indent=1)


if __name__ == "__main__":
    main()
