#!/usr/bin/env python
"""Offline (CPU, float64) model of a "region of interest" stage in front of pass 2 of calc_vert_volume: every z-line that
survives the row pass gets a conservative voxel range [a, b] that contains all of its voxels that can still be alive (voronoi and
edge planes: convex; wall wedges: subtracted), lines with an empty range are dropped, and pass 2 runs on [a, b] only -- its family
gating then sees the planes that bound the alive region instead of every plane that crosses the 41-voxel line.

    python tools/analyze_vv_roi.py [nverts] [dr] [margin_voxels]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from crixus_b200 import meshgen as mg  # noqa: E402
from oracle import orc  # noqa: E402
from analyze_vv_gating import fan_of  # noqa: E402

GS, GRES = 41, 20


def analyse(i, pos, norm, ep, dr, rim_index, margin):
    eps = dr / GRES * 1e-4
    gdr = dr / 2.0 / GRES
    tris = fan_of(i, ep, norm, pos)
    T = len(tris)
    pi = pos[i, :3].astype(np.float64)
    av = -sum(norm[s, :3].astype(np.float64) for _, _, s in tris)
    g = (np.arange(GS) - GRES) * gdr
    GX, GY, GZ = np.meshgrid(g, g, g, indexing="ij")
    fV, fE, W0, W1, W2 = [], [], [], [], []
    for va_i, vb_i, s in tris:
        n = norm[s, :3].astype(np.float64)
        e1, e2 = pos[va_i, :3] - pi, pos[vb_i, :3] - pi
        c9, c10, c11 = np.cross(e1, e2), np.cross(e2, av), np.cross(av, e1)
        if n.dot(c9) > 0:
            c9, c10, c11 = -c9, -c10, -c11
        adj = rim_index.get((min(va_i, vb_i), max(va_i, vb_i)), [])
        en = np.zeros(3)
        for sj in adj[:2]:
            en = en + norm[sj, :3]
        en = np.cross(en, pos[va_i, :3] - pos[vb_i, :3]) if adj else en
        if en.dot(e1) < 0:
            en = -en
        lin = lambda vec, off: vec[0] * (GX - off[0]) + vec[1] * (GY - off[1]) + vec[2] * (GZ - off[2])  # noqa: E731
        fV.append(lin(e1, e1 / 2)); fE.append(lin(en, e1))
        W0.append(lin(c9, np.zeros(3))); W1.append(lin(c10, np.zeros(3))); W2.append(lin(c11, np.zeros(3)))
    fV, fE, W0, W1, W2 = map(np.array, (fV, fE, W0, W1, W2))      # (T, 41, 41, 41)
    killV, killE = fV > eps, fE > eps
    inside = (W0 >= -eps) & (W1 >= -eps) & (W2 >= -eps)
    killW = inside & (W0 >= eps)
    # row pass: a line dies if ONE plane / wedge removes all of it
    dead = (killV.all(3) | killE.all(3) | killW.all(3)).any(0)
    alive = ~(killV.any(0) | killE.any(0) | killW.any(0))       # (41,41,41) the truth
    surv = ~dead
    line_alive = alive.any(2)
    # ROI of a line: first / last alive voxel, grown by the margin
    idx = np.arange(GS)
    first = np.where(alive, idx, GS).min(2)
    last = np.where(alive, idx, -1).max(2)
    a = np.clip(first - margin, 0, GS - 1)
    b = np.clip(last + margin, 0, GS - 1)
    roi_lines = surv & line_alive
    res = {"T": T, "surv": surv.sum(), "alive_lines": roi_lines.sum(), "roi_len": (b - a + 1)[roi_lines].sum()}
    # gating of pass 2, batches of 32 lines in index order: (1) today: all surviving lines, ends 0 / 40; (2) ROI lines, ends a / b
    K, L = np.nonzero(surv)
    K2, L2 = np.nonzero(roi_lines)
    neps = -eps

    def gate(Ks, Ls, A, B, tag):
        nb = (Ks.size + 31) // 32
        ent = {"V": 0, "E": 0, "W": 0}
        lanes = {"V": 0, "E": 0, "W": 0}
        srch = {"V": 0, "E": 0, "W": 0}
        for bi in range(nb):
            k, l = Ks[bi * 32:bi * 32 + 32], Ls[bi * 32:bi * 32 + 32]
            ea, eb = A[k, l], B[k, l]
            for t in range(T):
                for f, name in ((fV, "V"), (fE, "E")):
                    va, vb = f[t, k, l, ea], f[t, k, l, eb]
                    touch = ~((va <= neps) & (vb <= neps))
                    ent[name] += touch.any(); lanes[name] += touch.sum()
                    srch[name] += (((va > eps) != (vb > eps)) | ((va > neps) != (vb > neps))).any()
                w = [(Wx[t, k, l, ea], Wx[t, k, l, eb]) for Wx in (W0, W1, W2)]
                none = np.zeros(k.size, dtype=bool)
                for va, vb in w:
                    none |= (va < neps) & (vb < neps)
                ent["W"] += (~none).any(); lanes["W"] += (~none).sum()
                srch["W"] += sum((((va >= neps) != (vb >= neps)) & ~none).any() for va, vb in w)
        for n_ in "VEW":
            res["%s_ent_%s" % (tag, n_)] = ent[n_]
            res["%s_lanes_%s" % (tag, n_)] = lanes[n_]
            res["%s_srch_%s" % (tag, n_)] = srch[n_]
        res[tag + "_iters"] = nb * T
        res[tag + "_batches"] = nb
    z0 = np.zeros((GS, GS), dtype=np.int64)
    gate(K, L, z0, z0 + GS - 1, "now")
    gate(K2, L2, a, b, "roi")
    return res


def main():
    nverts = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    dr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.045
    margin = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    v, t, n = mg.spheric2(dr=dr, hfac=0.625)
    O = orc.Oracle()
    o = orc.run_pipeline(O, v[t], n, np.float32(dr), swap_normals=True, stages=())
    pos, norm, ep, nv = o["pos"].astype(np.float64), o["norm"], o["ep"], o["nvert"]
    rim = {}
    for s in range(ep.shape[0]):
        e = ep[s, :3]
        for a, b in ((e[0], e[1]), (e[1], e[2]), (e[2], e[0])):
            rim.setdefault((int(min(a, b)), int(max(a, b))), []).append(s)
    rng = np.random.default_rng(1)
    acc = {}
    for i in rng.choice(nv, nverts, replace=False):
        for k, val in analyse(int(i), pos, norm, ep, dr, rim, margin).items():
            acc[k] = acc.get(k, 0) + float(val)
    m = {k: val / nverts for k, val in acc.items()}
    print("mesh: spheric2 dr=%g h=0.625dr; %d sampled vertices; per-vertex means; ROI margin %d voxels" % (dr, nverts, margin))
    print("T = %.2f; lines surviving the row pass %.0f; lines with an alive voxel %.0f (mean ROI length %.1f voxels)"
          % (m["T"], m["surv"], m["alive_lines"], m["roi_len"] / max(m["alive_lines"], 1)))
    for tag, label in (("now", "today (all survivors, ends 0/40)"), ("roi", "ROI lines, ends a/b")):
        print("%-34s batches %5.1f iterations %6.1f | entered V %6.1f E %6.1f W %6.1f | searches V %6.1f E %6.1f W(planes) %6.1f | lanes V %.0f E %.0f W %.0f"
              % (label, m[tag + "_batches"], m[tag + "_iters"], m[tag + "_ent_V"], m[tag + "_ent_E"], m[tag + "_ent_W"],
                 m[tag + "_srch_V"], m[tag + "_srch_E"], m[tag + "_srch_W"], m[tag + "_lanes_V"], m[tag + "_lanes_E"], m[tag + "_lanes_W"]))


if __name__ == "__main__":
    main()
