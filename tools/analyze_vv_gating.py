#!/usr/bin/env python
"""Offline (CPU, float64) statistics of calc_vert_volume's pass 2 on a config-2-like mesh: how many (batch of 32 surviving
z-lines, fan triangle, plane family) iterations does warp-level gating execute, against what the lanes really need, for the
current batching (survivors in line-index order) and for sector-coherent batching (survivors sorted by their angle around the
vertex normal)?  Planning aid for the next optimisation of k_vert_volume; not part of the product.

    python tools/analyze_vv_gating.py [nverts] [dr]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from crixus_b200 import meshgen as mg  # noqa: E402
from oracle import orc  # noqa: E402

GS, GRES = 41, 20


def fan_of(i, ep, norm, pos):
    inc = np.nonzero((ep[:, :3] == i).any(1))[0]
    tris = []
    for s in inc:
        e = ep[s, :3]
        c = int(np.nonzero(e == i)[0][0])
        tris.append((int(e[(c + 1) % 3]), int(e[(c + 2) % 3]), int(s)))
    return tris


def analyse(i, pos, norm, ep, dr, rim_index):
    eps = dr / GRES * 1e-4
    gdr = dr / 2.0 / GRES
    tris = fan_of(i, ep, norm, pos)
    T = len(tris)
    pi = pos[i, :3].astype(np.float64)
    av = -sum(norm[s, :3].astype(np.float64) for _, _, s in tris)
    g = (np.arange(GS) - GRES) * gdr
    GX, GY = np.meshgrid(g, g, indexing="ij")              # line (k, l): x = g[k], y = g[l]
    zA, zB = g[0], g[-1]
    touchV, touchE, touchW = [], [], []
    in01, in012, killW = [], [], []          # per line: some voxel inside planes 0&1 / inside the tetrahedron / removed by it
    forms = []                               # per triangle: linear forms (vector, constant) of sV, sE, w0, w1, w2
    G3 = g[None, None, :]
    dead = np.zeros((GS, GS), dtype=bool)
    cutV, cutE, cutW = [], [], []
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

        def ends(vec, off):          # vec.(offset - off) at both ends of every line
            base = vec[0] * (GX - off[0]) + vec[1] * (GY - off[1])
            return base + vec[2] * (zA - off[2]), base + vec[2] * (zB - off[2])
        sVa, sVb = ends(e1, e1 / 2)
        sEa, sEb = ends(en, e1)
        w0a, w0b = ends(c9, np.zeros(3))
        w1a, w1b = ends(c10, np.zeros(3))
        w2a, w2b = ends(c11, np.zeros(3))
        touchV.append(~((sVa <= -eps) & (sVb <= -eps)))
        touchE.append(~((sEa <= -eps) & (sEb <= -eps)))
        noneW = ((w0a < -eps) & (w0b < -eps)) | ((w1a < -eps) & (w1b < -eps)) | ((w2a < -eps) & (w2b < -eps))
        touchW.append(~noneW)
        cutV.append((sVa > eps) != (sVb > eps))
        cutE.append((sEa > eps) != (sEb > eps))
        cutW.append(((w0a >= -eps) != (w0b >= -eps)) | ((w1a >= -eps) != (w1b >= -eps)) | ((w2a >= -eps) != (w2b >= -eps)))
        w0 = c9[0] * GX[:, :, None] + c9[1] * GY[:, :, None] + c9[2] * G3
        w1 = c10[0] * GX[:, :, None] + c10[1] * GY[:, :, None] + c10[2] * G3
        w2 = c11[0] * GX[:, :, None] + c11[1] * GY[:, :, None] + c11[2] * G3
        forms.append(((e1, -e1.dot(e1) / 2), (en, -en.dot(e1)), (c9, 0.0), (c10, 0.0), (c11, 0.0)))
        i01 = (w0 >= -eps) & (w1 >= -eps)
        in01.append(i01.any(2))
        in012.append((i01 & (w2 >= -eps)).any(2))
        killW.append((i01 & (w2 >= -eps) & (w0 >= eps)).any(2))
        dead |= ((sVa > eps) & (sVb > eps)) | ((sEa > eps) & (sEb > eps))
        dead |= (w0a >= eps) & (w0b >= eps) & (w1a >= -eps) & (w1b >= -eps) & (w2a >= -eps) & (w2b >= -eps)
    surv = np.nonzero(~dead.ravel())[0]                      # line index = k*41 + l, the kernel's order
    fam = {"V": np.array(touchV).reshape(T, -1)[:, surv], "E": np.array(touchE).reshape(T, -1)[:, surv],
           "W": np.array(touchW).reshape(T, -1)[:, surv]}
    cut = {"V": np.array(cutV).reshape(T, -1)[:, surv], "E": np.array(cutE).reshape(T, -1)[:, surv],
           "W": np.array(cutW).reshape(T, -1)[:, surv]}
    # sector key: angle of the line's (x, y) offset ... around the vertex normal projected to the voxel cube's x-y plane is not
    # meaningful for walls parallel to z, so use the index of the fan triangle whose voronoi cell the line's MIDPOINT lies in
    mid = np.stack([GX.ravel()[surv], GY.ravel()[surv], np.zeros(surv.size)], 1)
    dmin = np.full(surv.size, np.inf)
    key = np.zeros(surv.size, dtype=np.int64)
    for t, (va_i, vb_i, s) in enumerate(tris):
        c = (pos[va_i, :3] + pos[vb_i, :3]) / 3.0 - pi * (2.0 / 3.0)      # barycentre of the fan triangle, relative to pi
        d = ((mid - c) ** 2).sum(1)
        key = np.where(d < dmin, t, key)
        dmin = np.minimum(dmin, d)
    order_sector = np.lexsort((np.arange(surv.size), key))
    res = {"T": T, "survivors": surv.size}
    extra = {"in01": np.array(in01).reshape(T, -1)[:, surv], "in012": np.array(in012).reshape(T, -1)[:, surv],
             "killW": np.array(killW).reshape(T, -1)[:, surv]}
    for name, order in (("index", np.arange(surv.size)), ("sector", order_sector)):
        nb = (surv.size + 31) // 32
        pad = np.full(nb * 32, -1)
        pad[:surv.size] = order
        b = pad.reshape(nb, 32)
        for f in "VEW":
            m = np.concatenate([fam[f], np.zeros((T, 1), dtype=bool)], 1)[:, b]      # (T, nb, 32); -1 -> the padding column
            res["%s_entered_%s" % (name, f)] = int(m.any(2).sum())
            mc = np.concatenate([cut[f], np.zeros((T, 1), dtype=bool)], 1)[:, b]
            res["%s_search_%s" % (name, f)] = int(mc.any(2).sum())
        res[name + "_iters"] = nb * T
        # per-batch interval gating: bounding box of the batch's (x, y) offsets x {zA, zB}; a family of triangle t is skipped
        # iff the upper bound of its expression over the box is <= -eps (walls: of one of the three expressions)
        xs, ys = GX.ravel()[surv], GY.ravel()[surv]
        ent = {"V": 0, "E": 0, "W": 0}
        missed = 0
        for bi in range(nb):
            idx = b[bi][b[bi] >= 0]
            lo = np.array([xs[idx].min(), ys[idx].min(), zA]); hi = np.array([xs[idx].max(), ys[idx].max(), zB])
            for t_ in range(T):
                ub = [np.maximum(vec * lo, vec * hi).sum() + c0 for vec, c0 in forms[t_]]
                eV, eE, eW = ub[0] > -eps, ub[1] > -eps, min(ub[2], ub[3], ub[4]) >= -eps
                ent["V"] += eV; ent["E"] += eE; ent["W"] += eW
                missed += (fam["V"][t_, idx].any() and not eV) + (fam["E"][t_, idx].any() and not eE) + (fam["W"][t_, idx].any() and not eW)
        for f in "VEW":
            res["%s_bbox_%s" % (name, f)] = int(ent[f])
        res[name + "_bbox_missed"] = int(missed)
        mw = np.concatenate([fam["W"], np.zeros((T, 1), dtype=bool)], 1)[:, b].any(2)          # wall iterations entered
        for k2, arr in extra.items():
            mm = np.concatenate([arr & fam["W"], np.zeros((T, 1), dtype=bool)], 1)[:, b].any(2)
            res["%s_W_%s" % (name, k2)] = int((mw & mm).sum())
    for f in "VEW":
        res["lanes_touch_" + f] = int(fam[f].sum())
        res["lanes_cut_" + f] = int(cut[f].sum())
    return res


def main():
    nverts = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    dr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.045
    v, t, n = mg.spheric2(dr=dr, hfac=0.625)
    O = orc.Oracle()
    o = orc.run_pipeline(O, v[t], n, np.float32(dr), swap_normals=True, stages=())
    pos, norm, ep, nv = o["pos"], o["norm"], o["ep"], o["nvert"]
    rim = {}
    for s in range(ep.shape[0]):
        e = ep[s, :3]
        for a, b in ((e[0], e[1]), (e[1], e[2]), (e[2], e[0])):
            rim.setdefault((int(min(a, b)), int(max(a, b))), []).append(s)
    rng = np.random.default_rng(1)
    acc = {}
    ids = rng.choice(nv, nverts, replace=False)
    for i in ids:
        r = analyse(int(i), pos, norm, ep, dr, rim)
        for k, val in r.items():
            acc[k] = acc.get(k, 0) + val
    m = {k: val / nverts for k, val in acc.items()}
    print("mesh: spheric2 dr=%g h=0.625dr, %d triangles; %d sampled vertices; per-vertex means" % (dr, ep.shape[0], nverts))
    print("T = %.2f, surviving lines = %.0f of 1681, (batch, triangle) iterations = %.0f" % (m["T"], m["survivors"], m["index_iters"]))
    for f, label in (("V", "voronoi plane"), ("E", "edge plane"), ("W", "wall tetrahedron")):
        need = m["lanes_touch_" + f] / 32.0
        print("%-17s lanes touching %8.0f (= %.0f full warps) | entered: index order %6.0f, sector order %6.0f | with a search: index %6.0f, sector %6.0f, lanes cut %.0f"
              % (label, m["lanes_touch_" + f], need, m["index_entered_" + f], m["sector_entered_" + f], m["index_search_" + f],
                 m["sector_search_" + f], m["lanes_cut_" + f]))
    report_walls(m)
    print("per-batch interval gating (bounding box of the batch, one lane per triangle): entered V %.0f E %.0f W %.0f (exact per-lane "
          "gating: %.0f %.0f %.0f); blocks it would wrongly skip: %g" % (m["index_bbox_V"], m["index_bbox_E"], m["index_bbox_W"],
          m["index_entered_V"], m["index_entered_E"], m["index_entered_W"], m["index_bbox_missed"]))


def report_walls(m):
    print("wall iterations entered %.0f: some lane has a voxel inside planes 0 and 1: %.0f, inside all three: %.0f, removes a voxel: %.0f"
          % (m["index_entered_W"], m["index_W_in01"], m["index_W_in012"], m["index_W_killW"]))


if __name__ == "__main__":
    main()
