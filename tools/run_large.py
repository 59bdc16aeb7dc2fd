#!/usr/bin/env python
"""Runs one of the large BASELINE.json configurations on one B200 (device weld + hot path), times the stages with CUDA
events and checks size-independent properties against the CPU oracle on samples:

    gpurun -- python tools/run_large.py channel 1581          # configs[2]: ~1e7 triangles, x/y periodic
    gpurun -- python tools/run_large.py sphere 1000           # configs[3]: ~4.8e7 triangles, 1001^3 lattice
    gpurun -- python tools/run_large.py tank 2154             # configs[4]: ~2.2e8 triangles, ~1e10-cell lattice (64-bit indices)

Writes gpurun_out/large_<name>_<size>.json.  Checks: device weld vertex count against the closed-form count of the
generator; sampled vertex ranges bit-exact against the oracle; fill popcount = running count; fill idempotence; closure
(sampled unset cells next to set cells are blocked according to the oracle's checkCollision)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from crixus_b200 import api, meshgen as mg  # noqa: E402
from crixus_b200.dist import ShardedHotPath  # noqa: E402
from oracle import orc  # noqa: E402


def _op(p):
    q = orc.Params()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
    return q


def build(kind, size, mesh=True):
    v = t = n = None
    if kind == "channel":
        dr = 0.01
        H = max(6, size // 4) * 1.25 * dr
        if mesh:
            v, t, n = mg.channel(n=size, dr=dr)
            H = float(v[:, 2].max())
        L = size * 1.25 * dr
        w = dict(dr=np.float32(dr), swap_normals=True, periodic=(True, True, False), fills=[("geometry", (L / 2, L / 2, H / 2), np.float32(dr))])
    elif kind == "sphere":
        dr = 1.0 / size
        if mesh:
            v, t, n = mg.sphere_in_tank(side_cells=size, dr=dr, hfac=0.5, sphere_levels=7 if size >= 400 else 5)
        w = dict(dr=np.float32(dr), swap_normals=True, periodic=(False, False, False), fills=[("geometry", (4 * dr, 4 * dr, 4 * dr), np.float32(dr))])
    elif kind == "tank":
        dr = 1.0 / size
        if mesh:
            v, t, n = mg.box_tank((0, 0, 0), (1, 1, 1), 0.53 * dr, inward=False)
        w = dict(dr=np.float32(dr), swap_normals=True, periodic=(False, False, False), fills=[("geometry", (0.5, 0.5, 0.5), np.float32(dr))])
    else:
        raise SystemExit("unknown configuration " + kind)
    w.update(zero_open=False, container=None, fshape=None, name="%s %d" % (kind, size))
    return (v, t, n, w) if mesh else w


def main():
    kind, size = sys.argv[1], int(sys.argv[2])
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:      # launched by torch.distributed.run: one process per GPU, every rank holds the (replicated) mesh
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"config": kind, "size": size, "n_gpus": world}
    ctx = api.Context(local)
    if world > 1:   # the library's own NCCL communicator (as bench.py): rank 0 creates the id, torch.distributed ships the 128 bytes
        box = [api.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.dist_init(box[0], world, rank)
    dev = ctx.device
    if rank == 0:      # rank 0 generates and welds the mesh; the other ranks receive the welded arrays over NCCL
        t0 = time.time()
        v, t, n, w = build(kind, size)
        out["mesh_gen_s"] = time.time() - t0
        nbe = t.shape[0]
        corners = np.ascontiguousarray(v[t], dtype=np.float32).reshape(-1, 3)
        # ---- device weld (timed twice: first call grows the scratch arena)
        dc = torch.from_numpy(corners).to(dev)
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.time()
            pos4, ep4, lo, hi = ctx.weld_device(dc, w["dr"])
            torch.cuda.synchronize()
            out["weld_device_s"] = time.time() - t0
        nv = int(pos4.shape[0])
        out.update(triangles=int(nbe), vertices=nv, generator_vertices=int(v.shape[0]))
        assert nv == v.shape[0], "device weld found %d vertices, the indexed generator mesh has %d" % (nv, v.shape[0])
        norm4 = np.zeros((nbe, 4), dtype=np.float32)
        norm4[:, :3] = n
        del dc, corners, v, t, n
        norm4_d = torch.from_numpy(norm4).to(dev) if world > 1 else None
    else:
        w = build(kind, size, mesh=False)
    if world > 1:
        hdr = torch.zeros(8, dtype=torch.float64, device=dev)
        if rank == 0:
            hdr[:] = torch.tensor([nbe, nv] + [float(x) for x in lo] + [float(x) for x in hi], dtype=torch.float64)
        dist.broadcast(hdr, 0)
        h = hdr.cpu().numpy()
        nbe, nv = int(h[0]), int(h[1])
        lo, hi = h[2:5].astype(np.float32), h[5:8].astype(np.float32)
        if rank != 0:
            pos4 = torch.empty((nv, 4), dtype=torch.float32, device=dev)
            ep4 = torch.empty((nbe, 4), dtype=torch.int32, device=dev)
            norm4_d = torch.empty((nbe, 4), dtype=torch.float32, device=dev)
        for x in (pos4, ep4, norm4_d):
            dist.broadcast(x, 0)
        if rank != 0:
            norm4 = norm4_d.cpu().numpy()
        del norm4_d
    w.update(nvert=nv, nbe=nbe, pos=pos4.cpu().numpy(), ep=ep4.cpu().numpy(), norm=norm4, bbmin=lo, bbmax=hi)
    del pos4, ep4
    torch.cuda.empty_cache()
    # ---- hot path, device resident
    hp = ShardedHotPath(ctx, w, rank, world, byte_stats=False)
    torch.cuda.synchronize()
    steps = 2
    if world > 1:
        dist.barrier()
    for _ in range(steps):
        hp.step(timed=True)
    torch.cuda.synchronize()
    ms = hp.collect_stage_ms() / steps          # max over ranks
    dimg = [int(x) for x in hp.dimg]
    G = dimg[0] * dimg[1] * dimg[2]
    out.update(stage_ms={k: float(x) for k, x in zip(hp.STAGES, ms)}, lattice=dimg, lattice_cells=G, nfluid=int(hp.nfluid),
               fill_rounds=int(hp.fill_rounds), alive_vertices=int(hp.nalive), pass_ms=float(ms.sum()),
               triangles_per_s=nbe / (ms.sum() * 1e-3), verts_per_s=hp.nalive / (ms[hp.STAGES.index("vert_volume")] * 1e-3),
               fill_cells_per_s=G / (ms[hp.STAGES.index("fill")] * 1e-3), hbm_gb_in_use=torch.cuda.memory_allocated() / 1e9)
    out["parallelism"] = hp.parallelism
    if rank == 0:
        print(json.dumps(out), flush=True)
    if rank != 0:       # the checks below run on rank 0 (every rank holds the gathered volumes and the merged bitfield)
        dist.barrier()
        ctx.dist_finalize()
        dist.destroy_process_group()
        return
    # ---- properties
    O = orc.Oracle()
    pos, norm, ep = hp.pos.cpu().numpy(), hp.norm.cpu().numpy(), hp.ep.cpu().numpy()
    cell_idx, trisize, vol = hp.cell_idx.cpu().numpy(), hp.trisize.cpu().numpy(), hp.vol.cpu().numpy()
    alive = pos[:nv, 0] > -1e9
    assert cell_idx[0] == 0 and cell_idx[-1] == nbe
    pv = hp.p.copy()
    for j in range(3):
        pv.per[j] = int(w["periodic"][j])
    rng = np.random.default_rng(3)
    checked = 0
    t0 = time.time()
    for s in sorted(rng.integers(0, nv - 100, 4)):
        a, b = int(s), int(s) + 100
        ref = O.calc_vert_volume(_op(pv), pos, norm, ep, trisize, cell_idx, nv, nbe, False, a, b)
        m = alive[a:b]
        assert np.array_equal(vol[a:b][m].view(np.uint32), ref[a:b][m].view(np.uint32)), "volumes differ from the oracle in [%d,%d)" % (a, b)
        checked += int(m.sum())
    out["volumes_checked_bit_exact"] = checked
    out["oracle_volume_s"] = time.time() - t0
    dr3 = float(w["dr"]) ** 3
    assert np.all(vol[alive] > 0) and np.all(vol[alive] <= 1.0001 * dr3)
    # fill: popcount, idempotence, closure on sampled cells
    fpos = hp.fpos.cpu().numpy().view(np.uint32)
    pop = int(np.bitwise_count(fpos).sum()) if hasattr(np, "bitwise_count") else int(np.unpackbits(fpos.view(np.uint8)).sum())
    assert pop == hp.nfluid, (pop, hp.nfluid)
    f = [x for x in w["fills"] if x[0] == "geometry"][0]
    p = hp.pf.copy()
    p.dr_wall = np.float32(f[2])
    seed = api.seed_index(p, f[1])
    f1 = hp.fpos.clone()
    assert ctx.fill_geometry(p, f1, hp.norm, hp.ep, hp.pos, nbe, 0, hp.cell_idx, seed) == 0 and torch.equal(f1, hp.fpos)
    out["fill_idempotent"] = True      # at N > 1 this is also "the sharded fill equals the single-GPU fixed point"

    def bit(i, j, k):
        idx = i + dimg[0] * (j + dimg[1] * k)
        return (fpos[idx >> 5] >> np.uint32(idx & 31)) & 1
    # cells near the walls are the interesting ones: sample in the boundary layers of the lattice and around the seed
    pairs = []
    tries = 0
    want_pairs = int(os.environ.get("CX_LARGE_CLOSURE_PAIRS", "600"))
    while len(pairs) < want_pairs and tries < 4_000_000:
        tries += 1
        ax = int(rng.integers(0, 3))
        c = [int(rng.integers(0, dimg[a])) for a in range(3)]
        c[ax] = int(rng.integers(0, min(6, dimg[ax]))) if rng.random() < 0.5 else dimg[ax] - 1 - int(rng.integers(0, min(6, dimg[ax])))
        d = int(rng.integers(0, 3))
        e = list(c)
        e[d] += 1 if rng.random() < 0.5 else -1
        if not (0 <= e[d] < dimg[d]):
            continue
        bs, be = bit(*c), bit(*e)
        if bs == 0 and be == 1:
            pairs.append((tuple(c), tuple(e)))
    for s, e in pairs:
        assert O.check_collision(_op(p), s, e, norm, ep, pos, nbe, 0, cell_idx), "cell %r should have been filled from %r" % (s, e)
    out["closure_pairs_checked"] = len(pairs)
    print(json.dumps(out), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "large_%s_%d%s.json" % (kind, size, "" if world == 1 else "_n%d" % world)), "w"), indent=1)
    if world > 1:
        dist.barrier()
        ctx.dist_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
