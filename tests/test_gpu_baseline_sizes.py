"""BASELINE.json configs[2] and configs[3] at their FULL sizes (1e7 triangles / periodic; 4.8e7 triangles / 1e9-cell lattice) and
the whole-file comparison with the reference CUDA binary at the largest size it survives, inside the driver-run `-m gpu` suite.

At these sizes the oracle cannot run whole; what is checked:
  * the device weld finds exactly the generator's vertex count;
  * sampled vertex ranges bit-exact against the CPU oracle (>= 1000 vertices per configuration, >= 10 000 at configs[1]);
  * fill: popcount == running count, idempotence (a second fill adds nothing: the result is a fixed point), closure on sampled
    (unset cell, set neighbour) pairs through the oracle's checkCollision (it is the LEAST fixed point's complement that is
    blocked), and the multi-GPU form -- two in-process ranks, z-slab floods with halo exchange -- reproduces the bitfield;
  * 763 776-triangle SPHERIC-2 level: our CLI's .vtu against the reference binary's, particle by particle, bit for bit.
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

from crixus_b200 import api, workloads as wl
from crixus_b200.dist import ShardedHotPath
from test_gpu_dist import _run_ranks

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _op(p):
    from oracle import orc
    q = orc.Params()
    assert C.sizeof(q) == C.sizeof(p)
    C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
    return q


@pytest.fixture(scope="module")
def ctx():
    return api.Context(0)


def _device_workload(ctx, cfg, size=None):
    """the workload the way bench.py sets it up: generated on the host, welded on the device"""
    name, opts, gen = wl.bench_config(cfg, size)
    corners, normals = gen()
    nbe = corners.shape[0]
    dev = ctx.device
    dcorn = torch.from_numpy(corners.reshape(-1, 3)).to(dev)
    expected_nvert = None
    pos4, ep4, lo, hi = ctx.weld_device(dcorn, opts["dr"])
    norm4 = torch.zeros((nbe, 4), dtype=torch.float32, device=dev)
    norm4[:, :3] = torch.from_numpy(normals).to(dev)
    w = dict(opts, name=name, nbe=nbe, nvert=int(pos4.shape[0]), pos=pos4.clone(), ep=ep4, norm=norm4, bbmin=lo, bbmax=hi)
    del dcorn, pos4, corners, normals
    torch.cuda.empty_cache()
    return w, expected_nvert


def _check_large(hp, oracle, nranges=8, range_len=150, npairs=400, dist_ranks=2):
    ctx, w, nv, nbe = hp.ctx, hp.w, hp.nv, hp.nbe
    pos, norm, ep = hp.pos.cpu().numpy(), hp.norm.cpu().numpy(), hp.ep.cpu().numpy()
    cell_idx, trisize, vol = hp.cell_idx.cpu().numpy(), hp.trisize.cpu().numpy(), hp.vol.cpu().numpy()
    alive = pos[:nv, 0] > -1e9
    assert cell_idx[0] == 0 and cell_idx[-1] == nbe
    pv = hp.p.copy()
    for j in range(3):
        pv.per[j] = int(w["periodic"][j])
    rng = np.random.default_rng(5)
    checked = 0
    for s in sorted(rng.integers(0, nv - range_len, nranges)):
        a, b = int(s), int(s) + range_len
        ref = oracle.calc_vert_volume(_op(pv), pos, norm, ep, trisize, cell_idx, nv, nbe, False, a, b)
        m = alive[a:b]
        np.testing.assert_array_equal(vol[a:b][m].view(np.uint32), ref[a:b][m].view(np.uint32), err_msg="volumes in [%d,%d)" % (a, b))
        checked += int(m.sum())
    assert checked >= 0.9 * nranges * range_len
    dr3 = float(w["dr"]) ** 3
    assert np.all(vol[alive] > 0) and np.all(vol[alive] <= 1.0001 * dr3)
    # ---- fill
    dimg = [int(x) for x in hp.dimg]
    fpos = hp.fpos.cpu().numpy().view(np.uint32)
    pop = int(np.bitwise_count(fpos).sum())
    assert pop == hp.nfluid and 0 < pop < dimg[0] * dimg[1] * dimg[2]
    f = [x for x in w["fills"] if x[0] == "geometry"][0]
    p = hp.pf.copy()
    p.dr_wall = np.float32(f[2])
    seed = api.seed_index(p, f[1])
    f1 = hp.fpos.clone()
    assert ctx.fill_geometry(p, f1, hp.norm, hp.ep, hp.pos, nbe, 0, hp.cell_idx, seed) == 0 and torch.equal(f1, hp.fpos), "not a fixed point"
    del f1

    def bit(i, j, k):
        idx = i + dimg[0] * (j + dimg[1] * k)
        return (fpos[idx >> 5] >> np.uint32(idx & 31)) & 1
    pairs, tries = [], 0
    while len(pairs) < npairs and tries < 4_000_000:   # cells near the walls are the interesting ones
        tries += 1
        ax = int(rng.integers(0, 3))
        c = [int(rng.integers(0, dimg[a])) for a in range(3)]
        c[ax] = int(rng.integers(0, min(6, dimg[ax]))) if rng.random() < 0.5 else dimg[ax] - 1 - int(rng.integers(0, min(6, dimg[ax])))
        d = int(rng.integers(0, 3))
        e = list(c)
        e[d] += 1 if rng.random() < 0.5 else -1
        if not (0 <= e[d] < dimg[d]):
            continue
        if bit(*c) == 0 and bit(*e) == 1:
            pairs.append((tuple(c), tuple(e)))
    assert len(pairs) >= npairs // 2
    for s, e in pairs:
        assert oracle.check_collision(_op(p), s, e, norm, ep, pos, nbe, 0, cell_idx), "cell %r should have been filled from %r" % (s, e)
    # ---- the multi-GPU form at full size: in-process ranks, z-slab floods + halo exchange (cx_dist_fill_geometry, mode 2)
    want = ctx.hash_u32(hp.fpos)
    D = dict(pos=hp.pos, norm=hp.norm, ep=hp.ep, cell_idx=hp.cell_idx)
    nwords = hp.fpos.numel()

    def rank_fn(r, c2):
        c2.dist_set_fill_mode(2)
        fp = torch.zeros(nwords, dtype=torch.int32, device=hp.pos.device)
        torch.cuda.synchronize()
        n = c2.dist_fill_geometry(p, fp, D["norm"], D["ep"], D["pos"], nbe, 0, D["cell_idx"], seed)
        return n, c2.hash_u32(fp)
    for n, h in _run_ranks(dist_ranks, rank_fn):
        assert n == hp.nfluid and h == want, "sharded fill differs from the single-GPU fill"


def test_config3_channel_full_size(ctx, oracle):
    """BASELINE configs[2]: x/y-periodic channel, 9 998 244 triangles (1581 x 1581 quads per wall), periodicity_links at scale"""
    w, _ = _device_workload(ctx, 3)
    assert w["nbe"] == 2 * 2 * 1581 * 1581 and w["nvert"] == 2 * 1582 * 1582
    hp = ShardedHotPath(ctx, w, 0, 1, byte_stats=False)
    torch.cuda.synchronize()
    assert hp.nalive == 2 * 1581 * 1581           # the max-face copies are deleted and relinked
    ep = hp.ep[:, :3]
    assert bool((hp.pos[:hp.nv, 0][ep.reshape(-1).long()] > -1e9).all()), "a segment refers to a deleted vertex"
    _check_large(hp, oracle)


def test_config4_sphere_tank_full_size(ctx, oracle):
    """BASELINE configs[3]: sphere in tank, 48.3 M triangles, 1001^3-cell lattice -- the default workload of bench.py"""
    w, _ = _device_workload(ctx, 4)
    assert w["nbe"] > 48_000_000
    hp = ShardedHotPath(ctx, w, 0, 1, byte_stats=False)
    torch.cuda.synchronize()
    assert [int(x) for x in hp.dimg] == [1001, 1001, 1001]
    _check_large(hp, oracle, npairs=300)


def test_config2_ten_thousand_volumes(ctx, oracle):
    """BASELINE configs[1] (3.06 M triangles): 10 000+ vertex volumes bit-exact against the oracle on all host cores"""
    w = wl.config2()
    hp = ShardedHotPath(ctx, w, 0, 1, byte_stats=False)
    torch.cuda.synchronize()
    nv, nbe = hp.nv, hp.nbe
    pos, norm, ep = hp.pos.cpu().numpy(), hp.norm.cpu().numpy(), hp.ep.cpu().numpy()
    cell_idx, trisize, vol = hp.cell_idx.cpu().numpy(), hp.trisize.cpu().numpy(), hp.vol.cpu().numpy()
    rng = np.random.default_rng(17)
    n = 0
    for s in sorted(rng.integers(0, nv - 700, 16)):
        a, b = int(s), int(s) + 700
        ref = oracle.calc_vert_volume(_op(hp.p), pos, norm, ep, trisize, cell_idx, nv, nbe, False, a, b)
        np.testing.assert_array_equal(vol[a:b].view(np.uint32), ref[a:b].view(np.uint32))
        n += b - a
    assert n >= 10_000


def test_cli_equals_reference_binary_at_764k_triangles():
    """The largest SPHERIC-2 level the reference CUDA binary survives (763 776 triangles, dr/2): both command lines on the same
    .ini + STL files, .vtu files compared particle by particle (segments in canonical order) -- bit-identical."""
    ref = os.path.join(ROOT, "oracle", "_ref", "Crixus_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/Crixus_ref not built (needs /root/reference at build time)")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import time_reference_cuda as trc
    with tempfile.TemporaryDirectory() as d:
        row = trc.compare_level(d, 2, parity=True, timing=False)
    assert isinstance(row.get("parity"), dict) and row["parity"]["bit_exact"], row
    assert row["parity"]["segments"] == 763776
