"""Full-size GPU tests (BASELINE.json configs[1] = the bench workload, plus larger members of the channel and
sphere-in-tank families) through size-independent properties, and against the CPU oracle on samples:

  * calc_vert_volume: sampled vertex ranges bit-exact vs the oracle; index-range sharding invariance;
  * geometry fill: z-slab sharding invariance (bit-exact flags and counts), idempotence (a second fill adds nothing),
    closure (every unset cell next to a set cell is separated from it by the reference's checkCollision -- evaluated
    by the oracle on a sample -- i.e. the result is the fixed point), popcount == running count;
  * binning: cell_idx is an exclusive prefix ending at nbe and every segment sits in the cell of its barycentre;
  * periodic links: every vertex of the max face is deleted and linked, no dangling index survives in ep;
  * special-boundary tagging: cell-binned lookup == the oracle's all-pairs loop on wall-sized and adversarial grids.
"""
import numpy as np
import pytest
import torch

from crixus_b200 import api, workloads as wl
from crixus_b200.dist import ShardedHotPath, split_range

pytestmark = pytest.mark.gpu


def _op(p):
    """the product's cx_params -> the oracle's params struct (same layout, tests/test_host.py checks it)"""
    import ctypes as C
    from oracle import orc
    q = orc.Params()
    assert C.sizeof(q) == C.sizeof(p)
    C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
    return q


@pytest.fixture(scope="module")
def ctx():
    return api.Context(0)


def _run(ctx, w):
    hp = ShardedHotPath(ctx, w, 0, 1)      # one untimed pass happens in the constructor
    torch.cuda.synchronize()
    return hp


def _fill_inputs(hp):
    if hp.cnbe:
        return hp.fnorm, hp.fep, hp.fposv, hp.nbe + hp.cnbe
    return hp.norm, hp.ep, hp.pos, hp.nbe


def _geometry_fill(hp):
    f = [x for x in hp.w["fills"] if x[0] == "geometry"][0]
    p = hp.pf.copy()
    p.dr_wall = np.float32(f[2] if f[2] is not None else hp.w["dr"])
    return p, api.seed_index(p, f[1])


def _check_common(hp, oracle, nsample_ranges=4, range_len=150):
    ctx, w, nv, nbe = hp.ctx, hp.w, hp.nv, hp.nbe
    pos, norm, ep, surf = (t.cpu().numpy() for t in (hp.pos, hp.norm, hp.ep, hp.surf))
    cell_idx, trisize, vol = hp.cell_idx.cpu().numpy(), hp.trisize.cpu().numpy(), hp.vol.cpu().numpy()
    p = hp.p
    # ---- binning
    assert cell_idx[0] == 0 and cell_idx[-1] == nbe and np.all(np.diff(cell_idx) >= 0)
    bary = pos[nv:, :3]
    g = [np.clip(((bary[:, j] - np.float32(p.dmin[j])) / np.float32(p.griddr[j])).astype(np.int64), 0, p.gridn[j] - 1) for j in range(3)]
    cell = (g[0] * p.gridn[1] + g[1]) * p.gridn[2] + g[2]
    assert np.all(np.diff(cell) >= 0), "segments are not sorted by 3dr-cell"
    np.testing.assert_array_equal(np.searchsorted(cell, np.arange(p.gridn[3] + 1), side="left"), cell_idx)
    assert np.all(surf > 0) and np.isfinite(surf).all()
    # ---- connectivity
    alive = pos[:nv, 0] > -1e9
    np.testing.assert_array_equal(np.bincount(ep[:, :3].ravel(), minlength=nv), trisize)
    assert alive[ep[:, :3]].all(), "a segment refers to a deleted vertex"
    # ---- volumes: sampled ranges vs the oracle (bit-exact), sharding invariance, sanity
    pv = hp.p.copy()
    for j in range(3):
        pv.per[j] = int(w["periodic"][j])
    rng = np.random.default_rng(7)
    for s in sorted(rng.integers(0, max(nv - range_len, 1), nsample_ranges)):
        a, b = int(s), int(min(s + range_len, nv))
        ref = oracle.calc_vert_volume(_op(pv), pos, norm, ep, trisize, cell_idx, nv, nbe, w["zero_open"], a, b)
        m = alive[a:b]
        np.testing.assert_array_equal(vol[a:b][m].view(np.uint32), ref[a:b][m].view(np.uint32))
    vol2 = torch.zeros_like(hp.vol)
    for r in range(3):
        a, b = split_range(nv, 3, r)
        ctx.calc_vert_volume(pv, hp.pos, hp.norm, hp.ep, vol2, hp.trisize, hp.cell_idx, nv, nbe, w["zero_open"], a, b)
    al = torch.from_numpy(alive).to(vol2.device)
    assert torch.equal(vol2[al], hp.vol[al])
    # interleaved chunks (the multi-GPU sharding): 3 parts into zeroed arrays, summed
    tot = torch.zeros_like(hp.vol)
    for r in range(3):
        part = torch.zeros_like(hp.vol)
        ctx.calc_vert_volume_part(pv, hp.pos, hp.norm, hp.ep, part, hp.trisize, hp.cell_idx, nv, nbe, w["zero_open"], 1000, 3, r)
        own = ((torch.arange(nv, device=part.device) // 1000) % 3) == r
        assert not part[~own].any(), "a part wrote outside its chunks"
        tot += part
    assert torch.equal(tot[al], hp.vol[al])
    dr3 = float(w["dr"]) ** 3
    assert np.all(vol[alive] > 0) and np.all(vol[alive] <= 1.0001 * dr3)


def _check_fill(hp, oracle, nsample=1500):
    ctx, w = hp.ctx, hp.w
    gn, ge, gp, fnbe = _fill_inputs(hp)
    p, seed = _geometry_fill(hp)
    dimg = [int(x) for x in hp.dimg]
    G = dimg[0] * dimg[1] * dimg[2]
    f0 = hp.fpos.clone()
    bits = np.unpackbits(f0.cpu().numpy().view(np.uint8), bitorder="little")[:G].reshape(dimg[2], dimg[1], dimg[0]).astype(bool)
    assert int(bits.sum()) == hp.nfluid
    assert 0 < hp.nfluid < G
    # ---- idempotence: filling again from the same seed adds nothing
    f1 = f0.clone()
    assert ctx.fill_geometry(p, f1, gn, ge, gp, fnbe, hp.cnbe, hp.cell_idx, seed) == 0
    assert torch.equal(f1, f0)
    # ---- z-slab sharding on one GPU: 3 slabs, each with its own context and bitfield copy, halo planes exchanged by hand
    n = 3
    ctxs = [api.Context(0) for _ in range(n)]
    fs = [torch.zeros_like(f0) for _ in range(n)]
    if any(x[0] == "box" for x in w["fills"]):
        for x in w["fills"]:
            if x[0] == "box":
                for r in range(n):
                    ctxs[r].fill_box(hp.pf, fs[r], x[1], x[2])
    ks = [split_range(dimg[2], n, r) for r in range(n)]
    total, first = sum(1 for _ in ()), True
    P = dimg[0] * dimg[1]
    for c in ctxs:
        c.fill_set_max_rounds(4)        # the pipelined form: at most 4 rounds per call, halo exchange in between
    for outer in range(100000):
        changed, unfinished = 0, 0
        for r in range(n):
            changed += ctxs[r].fill_geometry_slab(p, fs[r], gn, ge, gp, fnbe, hp.cnbe, hp.cell_idx, seed, ks[r][0], ks[r][1], first)
            unfinished += ctxs[r].fill_unfinished()
        first = False
        total += changed
        if changed == 0 and unfinished == 0:
            break
        for r in range(n - 1):          # exchange the boundary planes between slab r and r+1 (packed words, OR)
            for src, dst, k in ((r, r + 1, ks[r][1] - 1), (r + 1, r, ks[r + 1][0])):
                a, b = (k * P) // 32, ((k + 1) * P + 31) // 32
                fs[dst][a:b] |= fs[src][a:b]
    for c in ctxs:
        c.fill_set_max_rounds(0)
    merged = torch.zeros_like(f0)
    for r in range(n):
        part = fs[r].cpu().numpy().view(np.uint32).copy()
        pb = np.unpackbits(part.view(np.uint8), bitorder="little")
        keep = np.zeros_like(pb)
        keep[ks[r][0] * P: ks[r][1] * P] = 1
        merged |= torch.from_numpy(np.packbits(pb & keep, bitorder="little").view(np.int32).copy()).to(merged.device)
    assert torch.equal(merged, f0), "z-slab sharded fill differs from the unsharded fill"
    # ---- "sharded tests + replicated flood" (small-lattice multi-GPU path): every emulated rank resolves its slab, the
    #      blocked planes are summed, each rank floods the whole lattice
    fr = [torch.zeros_like(f0) for _ in range(n)]
    for r in range(n):
        for x in w["fills"]:
            if x[0] == "box":
                ctxs[r].fill_box(hp.pf, fr[r], x[1], x[2])
    # the blocked planes are zero-copy views of context-owned memory: clone before the context is used again
    Bp = [ctxs[r].fill_resolve_part(p, fr[r], gn, ge, gp, fnbe, hp.cnbe, hp.cell_idx, n, r).clone() for r in range(n)]   # interleaved
    Bs = [ctxs[r].fill_resolve_slab(p, fr[r], gn, ge, gp, fnbe, hp.cnbe, hp.cell_idx, ks[r][0], ks[r][1]) for r in range(n)]  # z-slabs
    tot = Bs[0].clone()
    for r in range(1, n):
        tot += Bs[r]
    totp = Bp[0].clone()
    for r in range(1, n):
        totp += Bp[r]
    assert torch.equal(tot, totp), "interleaved and z-slab sharding of the directed tests disagree"
    for r in range(n):
        Bs[r].copy_(tot)
    for r in range(n):
        before = int(np.unpackbits(fr[r].cpu().numpy().view(np.uint8)).sum())
        got = ctxs[r].fill_propagate(p, fr[r], gn, ge, gp, fnbe, hp.cnbe, hp.cell_idx, seed)
        assert torch.equal(fr[r], f0), "sharded-tests/replicated-flood result differs from the unsharded fill"
        assert got == int(bits.sum()) - before
    # ---- large-lattice multi-GPU path: directed tests dealt to the ranks, planes merged, flood by z-slab (first_call = 2)
    fz = [torch.zeros_like(f0) for _ in range(n)]
    for r in range(n):
        for x in w["fills"]:
            if x[0] == "box":
                ctxs[r].fill_box(hp.pf, fz[r], x[1], x[2])
    views = [ctxs[r].fill_resolve_part(p, fz[r], gn, ge, gp, fnbe, hp.cnbe, hp.cell_idx, n, r) for r in range(n)]
    totz = views[0].clone()
    for r in range(1, n):
        totz += views[r]
    assert torch.equal(totz, totp)
    for r in range(n):
        views[r].copy_(totz)
    for c in ctxs:
        c.fill_set_max_rounds(3)
    first, total2 = True, 0
    for outer in range(100000):
        changed, unfinished = 0, 0
        for r in range(n):
            changed += ctxs[r].fill_geometry_slab(p, fz[r], gn, ge, gp, fnbe, hp.cnbe, hp.cell_idx, seed, ks[r][0], ks[r][1], 2 if first else 0)
            unfinished += ctxs[r].fill_unfinished()
        first = False
        total2 += changed
        if changed == 0 and unfinished == 0:
            break
        for r in range(n - 1):
            for src, dst, k in ((r, r + 1, ks[r][1] - 1), (r + 1, r, ks[r + 1][0])):
                a, b = (k * P) // 32, ((k + 1) * P + 31) // 32
                fz[dst][a:b] |= fz[src][a:b]
    assert total2 == total, (total2, total)
    merged2 = torch.zeros_like(f0)
    for r in range(n):
        pb = np.unpackbits(fz[r].cpu().numpy().view(np.uint8), bitorder="little")
        keep = np.zeros_like(pb)
        keep[ks[r][0] * P: ks[r][1] * P] = 1
        merged2 |= torch.from_numpy(np.packbits(pb & keep, bitorder="little").view(np.int32).copy()).to(merged2.device)
    assert torch.equal(merged2, f0), "dealt-tests / z-slab-flood result differs from the unsharded fill"
    for c in ctxs:
        c.fill_set_max_rounds(0)
        c.close()
    # ---- closure: unset cells next to set cells must be blocked by the reference's predicate (oracle on a sample)
    pos, norm, ep = (t.cpu().numpy() for t in (gp, gn, ge))
    cell_idx = hp.cell_idx.cpu().numpy()
    rng = np.random.default_rng(11)
    pairs = []
    for axis, (dz, dy, dx) in enumerate(((0, 0, 1), (0, 1, 0), (1, 0, 0))):
        a = bits[: dimg[2] - dz, : dimg[1] - dy, : dimg[0] - dx]
        b = bits[dz:, dy:, dx:]
        for s_is_low, m in ((True, ~a & b), (False, a & ~b)):
            idx = np.argwhere(m)
            if idx.shape[0] == 0:
                continue
            for k, j, i in idx[rng.choice(idx.shape[0], min(nsample // 6, idx.shape[0]), replace=False)]:
                lo, hi = (i, j, k), (i + dx, j + dy, k + dz)
                pairs.append((lo, hi) if s_is_low else (hi, lo))
    assert len(pairs) > 50
    for s, e in pairs:
        assert oracle.check_collision(_op(p), s, e, norm, ep, pos, fnbe, hp.cnbe, cell_idx), "cell %r should have been filled from %r" % (s, e)


def _check_special(hp, oracle, n=96):
    """special-boundary tagging at full size: two large inlet/outlet rectangles on the x-min / x-max walls plus an adversarial
    grid (cases.special_stress), cell-binned CUDA lookup against the oracle's all-pairs loop, bit-exact."""
    import cases
    from oracle.orc import f4
    ctx, w, nv, nbe = hp.ctx, hp.w, hp.nv, hp.nbe
    tris = np.asarray(w["tris"], dtype=np.float32)
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    y0, y1, z0, z1 = lo[1] + 0.2 * (hi[1] - lo[1]), lo[1] + 0.7 * (hi[1] - lo[1]), lo[2] + 0.1 * (hi[2] - lo[2]), lo[2] + 0.5 * (hi[2] - lo[2])
    walls = np.array([[[x, y0, z0], [x, y1, z0], [x, y1, z1]] for x in (lo[0], hi[0])] +
                     [[[x, y0, z0], [x, y1, z1], [x, y0, z1]] for x in (lo[0], hi[0])], dtype=np.float32)
    grids = [walls, cases.special_stress(tris, 3, n=n)]
    pos, ep, cell_idx = hp.pos.cpu().numpy(), hp.ep.cpu().numpy(), hp.cell_idx.cpu().numpy()
    got = torch.zeros(nv + nbe, dtype=torch.int32, device=hp.pos.device)
    ref = np.zeros(nv + nbe, dtype=np.int32)
    for k, g in enumerate(grids):
        sv, se = api.weld_host(g.reshape(-1, 3), np.float32(w["dr"]))
        sbpos, sbep = f4(sv), np.zeros((se.shape[0], 4), dtype=np.int32)
        sbep[:, :3] = se
        ctx.identify_special_boundary(hp.pf, hp.pos, hp.ep, hp.cell_idx, nv, nbe, torch.from_numpy(sbpos).to(got.device),
                                      torch.from_numpy(sbep).to(got.device), sbep.shape[0], got, k + 1)
        oracle.identify_special_boundary(_op(hp.pf), pos, ep, nv, nbe, sbpos, sbep, ref, k + 1)
    np.testing.assert_array_equal(got.cpu().numpy(), ref)
    assert (ref[nv:] == 1).sum() > 10 and (ref[nv:] == 2).sum() > 0


def test_config2_full_size(ctx, oracle):
    """BASELINE.json configs[1]: 3.06e6 triangles, 328 x 219 x 131 lattice, free-surface fill."""
    w = wl.config2(3)
    assert w["nbe"] > 3_000_000
    hp = _run(ctx, w)
    _check_common(hp, oracle)
    _check_fill(hp, oracle)
    _check_special(hp, oracle)


def test_config1(ctx, oracle):
    """BASELINE.json configs[0]: resources/spheric2.ini verbatim (box fill + geometry fill)."""
    hp = _run(ctx, wl.config1())
    _check_common(hp, oracle, nsample_ranges=6)
    _check_fill(hp, oracle)
    _check_special(hp, oracle, n=240)


def test_channel_family(ctx, oracle):
    """configs[2] family (x/y periodic channel) at 6.4e5 triangles: periodic pairing at scale."""
    w = wl.channel(n=400, dr=0.01)
    hp = _run(ctx, w)
    pos = hp.pos.cpu().numpy()
    nv = hp.nv
    deleted = pos[:nv, 0] < -1e9
    n = 400
    assert int(deleted.sum()) == 2 * (2 * (n + 1) - 1), "every vertex of the two max faces (floor and lid) must be linked and deleted"
    _check_common(hp, oracle)
    _check_fill(hp, oracle)


def test_sphere_tank_family(ctx, oracle):
    """configs[3] family (sphere in a closed tank): curved, non axis-aligned triangles, 100^3 lattice."""
    w = wl.sphere_tank(side_cells=100, dr=0.01, sphere_levels=5)
    hp = _run(ctx, w)
    _check_common(hp, oracle)
    _check_fill(hp, oracle)
    # the sphere's interior is never reached
    dimg = [int(x) for x in hp.dimg]
    bits = np.unpackbits(hp.fpos.cpu().numpy().view(np.uint8), bitorder="little")[: dimg[0] * dimg[1] * dimg[2]].reshape(dimg[2], dimg[1], dimg[0])
    c = [d // 2 for d in dimg]
    assert bits[c[2], c[1], c[0]] == 0
