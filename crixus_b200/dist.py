"""Multi-GPU harness above the C-ABI: one process per GPU.

The sharded algorithms live INSIDE the library (include/crixus_b200.h "multi-GPU": cx_dist_init, cx_dist_calc_vert_volume,
cx_dist_fill_geometry, cx_preprocess_dist -- NCCL collectives on the context's stream, csrc/dist.cu + csrc/fill.cu); this module
only ships the 128-byte NCCL id between the ranks (through torch.distributed, which the launcher already set up) and holds the
device-resident workload for bench.py.  What shards and what does not (DESIGN.md "multi-GPU"):
  * swap_normals / set_bound_elem / cell binning / calc_trisize / device weld are HBM-bound passes of ~100-150 B per segment; an
    all-gather of their outputs over NVLink would move more bytes than recomputing them, so every rank runs them on the full
    (replicated) mesh -- identical results on every rank because the binning order is stable.
  * calc_vert_volume (FP32-ALU bound, ~90 % of the time): interleaved 1024-vertex chunks, volumes all-reduced (4 B/vertex).
  * geometry fill: large lattices -- every rank resolves the directed tests of the z-slab it floods (no merge), slab floods
    with halo exchange, owned planes all-gathered; small lattices -- directed tests dealt by 3dr-cell chunk, blocked planes
    all-reduced, flood replicated.  Bit-exact across rank counts by the fixed-point property of the fill.

The pure partition helpers below restate the library's sharding rules; they are tested on CPU with gloo (world_size 2 and 3,
the CPU oracle as the compute backend, tests/test_dist.py) -- the same rules the C++ code follows (csrc/dist.h cx_split_range,
csrc/fill.cu chunk_owner)."""
from __future__ import annotations

import ctypes as C

import numpy as np


FILL_ROUNDS_PER_EXCHANGE = 8   # z-slab fill: propagation rounds between two halo exchanges (one tile layer per round)
VV_CHUNK = 1024     # vertices per chunk of the interleaved calc_vert_volume sharding

# lattices whose six blocked bit planes stay below this size use "sharded tests + replicated flood" on multi-GPU runs
SMALL_LATTICE_BYTES = 64 << 20


def split_range(n, world, rank):
    """contiguous index range [lo, hi) of rank; sizes differ by at most one"""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


FT_Y, FT_Z, DEAL_CHUNK = 16, 16, 64   # fill tiles (csrc/fill.cu FT_Y / FT_Z) and the chunk of the interleaved dealing


def dealt_chunk_owner(chunk, nparts):
    """Rank that resolves the directed tests of 3dr-cell chunk `chunk` (32 consecutive 3dr-cells; csrc/fill.cu chunk_owner):
    every group of nparts consecutive chunks is dealt one chunk per rank, the order rotated by a multiplicative hash of the
    group number.  (Plain round-robin is periodic in the cell grid: a wall is a plane of 3dr-cells, i.e. every gridn_z/32-th
    chunk, and could land on few ranks.)  Vectorised over numpy arrays."""
    chunk = np.asarray(chunk, dtype=np.int64)
    g, slot = chunk // nparts, chunk % nparts
    rot = ((g.astype(np.uint64) & np.uint64(0xFFFFFFFF)) * np.uint64(2654435761) & np.uint64(0xFFFFFFFF)) >> np.uint64(16)
    return ((slot - (rot % np.uint64(nparts)).astype(np.int64)) % nparts).astype(np.int64)


def cell3_chunk(gx, gy, gz, gridn):
    """chunk of the up-front resolve that holds 3dr-cell (gx, gy, gz): cells are numbered gx*ny*nz + gy*nz + gz, 32 per chunk"""
    return ((np.asarray(gx, dtype=np.int64) * gridn[1] + gy) * gridn[2] + gz) // 32


def split_weighted(weights, world, rank):
    """contiguous range [lo, hi) with ~equal sum of weights (vertex ranges balanced by trisize)"""
    c = np.concatenate([[0], np.cumsum(np.asarray(weights, dtype=np.int64))])
    tot = c[-1]
    lo = int(np.searchsorted(c, tot * rank / world, side="left"))
    hi = int(np.searchsorted(c, tot * (rank + 1) / world, side="left")) if rank + 1 < world else len(weights)
    return min(lo, len(weights)), min(max(hi, lo), len(weights))


def plane_words(dimg, k):
    """word range [w0, w1) of the packed reference bitfield that covers lattice plane k"""
    P = int(dimg[0]) * int(dimg[1])
    return (k * P) // 32, ((k + 1) * P + 31) // 32


def exchange_halos(fpos, dimg, k0, k1, rank, world, dist, device_ops):
    """Sends this rank's two boundary planes to its z-neighbours and ORs the received planes into fpos.
    fpos: int32 torch tensor holding the packed bitfield (CPU for gloo, CUDA for nccl)."""
    import torch
    kd = int(dimg[2])
    ops, recv = [], []
    if rank + 1 < world and k1 < kd:          # upper neighbour: send my top plane, receive its bottom plane (k1)
        a, b = plane_words(dimg, k1 - 1)
        ops.append(dist.P2POp(dist.isend, fpos[a:b].contiguous(), rank + 1))
        a, b = plane_words(dimg, k1)
        buf = torch.empty(b - a, dtype=fpos.dtype, device=fpos.device)
        ops.append(dist.P2POp(dist.irecv, buf, rank + 1))
        recv.append((a, b, buf))
    if rank > 0 and k0 > 0:                   # lower neighbour
        a, b = plane_words(dimg, k0)
        ops.append(dist.P2POp(dist.isend, fpos[a:b].contiguous(), rank - 1))
        a, b = plane_words(dimg, k0 - 1)
        buf = torch.empty(b - a, dtype=fpos.dtype, device=fpos.device)
        ops.append(dist.P2POp(dist.irecv, buf, rank - 1))
        recv.append((a, b, buf))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    for a, b, buf in recv:
        fpos[a:b] |= buf
    del device_ops


def distributed_fill(slab_fill, fpos, dimg, seed, rank, world, dist):
    """z-slab sharded geometry fill.  slab_fill(fpos, seed, k0, k1, first_call) -> newly set cells in the slab.
    Returns (total new cells over all ranks, outer iterations).  Afterwards every rank holds its own slab; call
    gather_slabs to assemble the full bitfield."""
    import torch
    k0, k1 = split_range(int(dimg[2]), world, rank)
    total, it = 0, 0
    while True:
        r = slab_fill(fpos, seed, k0, k1, it == 0) if k1 > k0 else 0
        n, unfinished = r if isinstance(r, tuple) else (r, 0)     # unfinished: stopped at its round limit with work left
        t = torch.tensor([n, unfinished], dtype=torch.int64, device=fpos.device)
        if world > 1:
            dist.all_reduce(t)
        it += 1
        tn, tu = (int(x) for x in t.tolist())
        total += tn
        if (tn == 0 and tu == 0) or world == 1:
            break
        exchange_halos(fpos, dimg, k0, k1, rank, world, dist, None)
    return total, it


def own_bits_only(fpos, dimg, k0, k1):
    """Clears every bit outside this rank's planes [k0, k1) (the received halo planes and everything it never owned)."""
    P = int(dimg[0]) * int(dimg[1])
    lo, hi = k0 * P, k1 * P
    w0, w1 = lo // 32, (hi + 31) // 32
    if hi <= lo:
        fpos.zero_()
        return fpos
    fpos[:w0] = 0
    fpos[w1:] = 0
    if lo % 32:
        m = (0xFFFFFFFF << (lo % 32)) & 0xFFFFFFFF
        fpos[w0] &= m - (1 << 32) if m >= (1 << 31) else m          # as a signed 32-bit value
    if hi % 32:
        m = (1 << (hi % 32)) - 1
        fpos[w1 - 1] &= m - (1 << 32) if m >= (1 << 31) else m
    return fpos


def gather_slabs(fpos, dimg, rank, world, dist):
    """Assembles the full bitfield on every rank.  NCCL has no bitwise-OR reduction, so each rank first clears the bits
    it does not own; the owned bit ranges are disjoint, hence the (wrapping) int32 SUM equals the OR."""
    if world > 1:
        k0, k1 = split_range(int(dimg[2]), world, rank)
        own_bits_only(fpos, dimg, k0, k1)
        dist.all_reduce(fpos, op=dist.ReduceOp.SUM)
    return fpos


# ---------------------------------------------------------------------------------------------------- GPU hot path
class ShardedHotPath:
    """Device-resident hot path for bench.py: holds the welded mesh in HBM and runs one pass per step()."""
    STAGES = ["restore", "set_bound_elem", "binning", "trisize", "vert_volume", "fill"]

    def __init__(self, ctx, w, rank=0, world=1, byte_stats=True):
        import torch
        from . import api
        self.ctx, self.w, self.rank, self.world = ctx, w, rank, world
        self.torch = torch
        dev = ctx.device
        nv, nbe = w["nvert"], w["nbe"]
        self.nv, self.nbe = nv, nbe
        # the welded mesh: numpy arrays (small workloads built on the host) or tensors already on the device (large ones, welded there)
        f = lambda a: a.to(dev) if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
        self.pos0, self.norm0, self.ep0 = f(w["pos"]), f(w["norm"]), f(w["ep"])
        self.p = api.params_domain(w["dr"], w["bbmin"], w["bbmax"])
        self.ncell3 = int(self.p.gridn[3])
        z4 = lambda n, dt=torch.float32: torch.zeros((n, 4), dtype=dt, device=dev)   # noqa: E731
        self.pre_pos, self.pos = z4(nv + nbe), z4(nv + nbe)
        self.pre_norm, self.norm = z4(nbe), z4(nbe)
        self.pre_ep, self.ep = z4(nbe, torch.int32), z4(nbe, torch.int32)
        self.pre_surf = torch.zeros(nbe, dtype=torch.float32, device=dev)
        self.surf = torch.zeros(nbe, dtype=torch.float32, device=dev)
        self.cell_idx = torch.zeros(self.ncell3 + 1, dtype=torch.int32, device=dev)
        self.trisize = torch.zeros(nv, dtype=torch.int32, device=dev)
        self.vol = torch.zeros(nv, dtype=torch.float32, device=dev)
        self.newlink = torch.zeros(nv, dtype=torch.int32, device=dev)
        # fill parameters (crixus.cu:464-467, 698-722)
        self.pf = self.p.copy()
        for j in range(3):
            self.pf.per[j] = int(w["periodic"][j])
        api.params_fill_eps(self.pf)
        if w["container"] is not None:
            api.params_container(self.pf, True, *w["container"])
        else:
            api.params_container(self.pf, False)
        self.dimg = api.fill_dims(self.pf)
        G = int(np.prod(self.dimg))
        self.fpos = torch.zeros((G + 31) // 32, dtype=torch.int32, device=dev)
        self.cnbe = 0
        if w["fshape"] is not None:
            fs_t, fs_n = w["fshape"]
            fv, fe = api.weld_host(np.ascontiguousarray(fs_t, dtype=np.float32).reshape(-1, 3), w["dr"])
            self.cnbe = fe.shape[0]
            fv4 = np.zeros((fv.shape[0], 4), dtype=np.float32); fv4[:, :3] = fv
            fn4 = np.zeros((self.cnbe, 4), dtype=np.float32); fn4[:, :3] = fs_n
            fe4 = np.zeros((self.cnbe, 4), dtype=np.int32); fe4[:, :3] = fe + nv
            self.fs_pos, self.fs_norm, self.fs_ep = f(fv4), f(fn4), f(fe4)
            self.fposv = z4(nv + fv.shape[0])
            self.fnorm, self.fep = z4(nbe + self.cnbe), z4(nbe + self.cnbe, torch.int32)
        self.v0, self.v1 = 0, nv
        self.k0, self.k1 = split_range(int(self.dimg[2]), world, rank)
        wr = (int(self.dimg[0]) + 31) // 32
        small = 6 * 4 * wr * int(self.dimg[1]) * int(self.dimg[2]) <= SMALL_LATTICE_BYTES or int(self.dimg[2]) < 2 * world
        self.parallelism = "1 GPU" if world == 1 else (
            "one process per GPU, NCCL inside the C-ABI (cx_dist_*); vert-volume: interleaved %d-vertex chunks over %d ranks + all-reduce; "
            "fill: %s; cheap stages and device weld replicated"
            % (VV_CHUNK, world, ("directed tests dealt to %d ranks by 3dr-cell chunk, blocked planes all-reduced, flood replicated" % world) if small else
               ("z-slabs: each of the %d ranks resolves the directed tests of the planes it floods (no merge), flood in %d z-slabs with a "
                "halo exchange over NCCL send/recv every %d rounds, owned planes all-gathered" % (world, world, FILL_ROUNDS_PER_EXCHANGE))))
        self.events, self.nfluid, self.fill_rounds = [], 0, 0
        self._sizes = None
        # one untimed pass to learn trisize (vertex-range balancing) and the S27 statistic for the byte count
        self.step()
        ts = self.trisize.cpu().numpy()
        alive = self.pos[:nv, 0].cpu().numpy() > -1e9
        self.nalive = int(alive.sum())
        self.mean_trisize = float(ts[alive].mean())
        own = ((np.arange(nv) // VV_CHUNK) % world) == rank
        self.local_verts = int((alive & own).sum())
        self.vv_bytes, self.vv_kernel_ms = 0.0, 1.0
        if not byte_stats:      # the S27 statistic below costs 27 passes over the 3dr-cell grid on the host (minutes at 1e10 cells)
            return
        cnt = np.diff(self.cell_idx.cpu().numpy()).reshape(self.p.gridn[0], self.p.gridn[1], self.p.gridn[2]).astype(np.int64)
        pad = np.pad(cnt, 1)
        s27 = sum(pad[1 + a:pad.shape[0] - 1 + a, 1 + b:pad.shape[1] - 1 + b, 1 + c:pad.shape[2] - 1 + c]
                  for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1))
        vpos = (w["pos"].cpu().numpy() if torch.is_tensor(w["pos"]) else w["pos"])[own, :3]
        g = [np.clip(((vpos[:, j] - np.float32(self.p.dmin[j])) / np.float32(self.p.griddr[j])).astype(np.int64), 0, self.p.gridn[j] - 1)
             for j in range(3)]
        S27 = s27[g[0], g[1], g[2]].astype(np.float64)
        self.vv_bytes = float((240.0 + 32.0 * S27 + 48.0 * ts[own]).sum())   # SURVEY.md §8d bytes / vertex
        self.vv_kernel_ms = 1.0

    # ---- one pass
    def step(self, timed=False):
        torch, ctx, w = self.torch, self.ctx, self.w
        nv, nbe = self.nv, self.nbe
        ev = []

        def mark():
            if timed:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                ev.append(e)
        mark()
        self.pre_pos[:nv].copy_(self.pos0)          # the welded mesh as uploaded by crixus.cu:295-297
        self.pre_norm.copy_(self.norm0)
        self.pre_ep.copy_(self.ep0)
        mark()
        if w["swap_normals"]:
            ctx.swap_normals(self.pre_norm)
        ctx.set_bound_elem(self.pre_pos, self.pre_norm, self.pre_surf, self.pre_ep, nbe, nv, want_zstat=False)
        mark()
        ctx.bin_segments(self.p, self.pre_pos, self.pos, self.pre_norm, self.norm, self.pre_ep, self.ep, self.pre_surf, self.surf,
                         self.cell_idx, nbe, nv)
        pv = self.p.copy()
        for idim in range(3):
            if w["periodic"][idim]:
                self.newlink.fill_(-1)
                ctx.find_links(self.p, self.pos, nv, self.newlink, idim)
                ctx.periodicity_links(self.ep, nbe, self.newlink)
            pv.per[idim] = int(w["periodic"][idim])
        mark()
        self.trisize.zero_()
        ctx.calc_trisize(self.ep, self.trisize, nbe)
        mark()
        if self.world == 1:
            ctx.calc_vert_volume(pv, self.pos, self.norm, self.ep, self.vol, self.trisize, self.cell_idx, nv, nbe, w["zero_open"], 0, nv)
        else:   # interleaved chunks + all-reduce inside the library (cx_dist_calc_vert_volume)
            ctx.dist_calc_vert_volume(pv, self.pos, self.norm, self.ep, self.vol, self.trisize, self.cell_idx, nv, nbe, w["zero_open"])
        mark()
        self._fill()
        mark()
        if timed:
            self.events.append(ev)
        return np.zeros(len(self.STAGES))

    def _fill(self):
        ctx, w, nv, nbe = self.ctx, self.w, self.nv, self.nbe
        from . import api
        self.fpos.zero_()
        if self.cnbe:
            self.fposv[:nv].copy_(self.pos[:nv]); self.fposv[nv:].copy_(self.fs_pos)
            self.fnorm[:nbe].copy_(self.norm); self.fnorm[nbe:].copy_(self.fs_norm)
            self.fep[:nbe].copy_(self.ep); self.fep[nbe:].copy_(self.fs_ep)
            gn, ge, gp, fnbe = self.fnorm, self.fep, self.fposv, nbe + self.cnbe
        else:
            gn, ge, gp, fnbe = self.norm, self.ep, self.pos, nbe
        nfl, rounds = 0, 0
        for f in w["fills"]:
            if f[0] == "box":
                nfl += ctx.fill_box(self.pf, self.fpos, f[1], f[2])
                continue
            p = self.pf.copy()
            p.dr_wall = np.float32(f[2] if f[2] is not None else w["dr"])
            seed = api.seed_index(p, f[1])
            if self.world == 1:
                nfl += ctx.fill_geometry(p, self.fpos, gn, ge, gp, fnbe, self.cnbe, self.cell_idx, seed)
            else:   # slab-owned directed tests, slab floods + halo exchange, all-gather: cx_dist_fill_geometry
                nfl += ctx.dist_fill_geometry(p, self.fpos, gn, ge, gp, fnbe, self.cnbe, self.cell_idx, seed)
            rounds += ctx.last_rounds
        self.nfluid, self.fill_rounds = int(nfl), int(rounds)

    def collect_stage_ms(self):
        torch = self.torch
        acc = np.zeros(len(self.STAGES))
        for ev in self.events:
            for s in range(len(self.STAGES)):
                acc[s] += ev[s].elapsed_time(ev[s + 1])
        self.events = []
        t = torch.tensor(acc, dtype=torch.float64, device=self.ctx.device)
        rank0 = acc.copy()
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        self._stage_rank_local = rank0
        return t.cpu().numpy()

    # ---- bitwise evidence
    def check(self):
        """popcount + order-independent 64-bit checksums (cx_hash_u32, computed on the device) of the fluid bitfield, the
        vertex-volume bits and the sorted geometry -- what a wrong bit anywhere would change"""
        hf, pop = self.ctx.hash_u32(self.fpos)
        hv, _ = self.ctx.hash_u32(self.vol)
        hg = sum(self.ctx.hash_u32(t)[0] for t in (self.pos, self.norm, self.ep, self.surf, self.cell_idx)) & 0xFFFFFFFFFFFFFFFF
        return {"nfluid_popcount": pop, "fpos_hash": "%016x" % hf, "vol_hash": "%016x" % hv, "geometry_hash": "%016x" % hg}

    def single_gpu_check(self):
        """the same inputs through the single-GPU calls (cx_calc_vert_volume, cx_fill_geometry) on this rank alone: the
        checksums a sharded run has to reproduce"""
        torch, ctx, w, nv, nbe = self.torch, self.ctx, self.w, self.nv, self.nbe
        from . import api
        pv = self.p.copy()
        for idim in range(3):
            pv.per[idim] = int(w["periodic"][idim])
        vol = torch.zeros_like(self.vol)
        ctx.calc_vert_volume(pv, self.pos, self.norm, self.ep, vol, self.trisize, self.cell_idx, nv, nbe, w["zero_open"], 0, nv)
        fpos = torch.zeros_like(self.fpos)
        gn, ge, gp, fnbe = (self.fnorm, self.fep, self.fposv, nbe + self.cnbe) if self.cnbe else (self.norm, self.ep, self.pos, nbe)
        for f in w["fills"]:
            if f[0] == "box":
                ctx.fill_box(self.pf, fpos, f[1], f[2])
                continue
            p = self.pf.copy()
            p.dr_wall = np.float32(f[2] if f[2] is not None else w["dr"])
            ctx.fill_geometry(p, fpos, gn, ge, gp, fnbe, self.cnbe, self.cell_idx, api.seed_index(p, f[1]))
        hf, pop = ctx.hash_u32(fpos)
        hv, _ = ctx.hash_u32(vol)
        return {"nfluid_popcount": pop, "fpos_hash": "%016x" % hf, "vol_hash": "%016x" % hv}

    # ---- end to end with host buffers
    def make_job(self, hcorn, hnorm3, keep):
        """cx_job over the raw STL content (corners + facet normals) in pinned host memory"""
        from . import api
        w, nbe = self.w, self.nbe
        job = api.CxJob()
        job.normals = C.cast(hnorm3.data_ptr(), C.POINTER(C.c_float))
        job.nbe = nbe
        job.corners = C.cast(hcorn.data_ptr(), C.POINTER(C.c_float))
        for j in range(3):
            job.periodic[j] = int(w["periodic"][j])
        job.dr = float(w["dr"])
        job.swap_normals, job.zero_open = int(w["swap_normals"]), int(w["zero_open"])
        if w["container"] is not None:
            job.use_container = 1
            for j in range(3):
                job.cmin[j], job.cmax[j] = w["container"][0][j], w["container"][1][j]
        if w["fshape"] is not None:
            fc = np.ascontiguousarray(w["fshape"][0], dtype=np.float32).reshape(-1)
            fn = np.ascontiguousarray(w["fshape"][1], dtype=np.float32).reshape(-1)
            keep += [fc, fn]
            job.fs_corners = fc.ctypes.data_as(C.POINTER(C.c_float))
            job.fs_normals = fn.ctypes.data_as(C.POINTER(C.c_float))
            job.fs_nbe = w["fshape"][0].shape[0]
        specs = (api.CxFillSpec * len(w["fills"]))()
        for i, f in enumerate(w["fills"]):
            if f[0] == "box":
                specs[i].option = 1
                for j in range(3):
                    specs[i].lo[j], specs[i].hi[j] = f[1][j], f[2][j]
            else:
                specs[i].option = 2
                for j in range(3):
                    specs[i].seed[j] = f[1][j]
                specs[i].dr_wall = float(f[2]) if f[2] is not None else 0.0
        job.fills, job.nfills = specs, len(w["fills"])
        keep.append(specs)
        return job

    def e2e(self, steps, barrier, hcorn, hnorm3):
        """The whole pass through the reference-facing C-ABI call on HOST buffers: cx_preprocess_host (one GPU) or
        cx_preprocess_dist (every rank uploads 1/N of the raw STL content, NCCL all-gather, sharded stages, every rank
        downloads its 1/N share of the results).  hcorn / hnorm3: pinned host tensors with the raw corners / facet normals."""
        import time
        torch, nv, nbe = self.torch, self.nv, self.nbe
        keep = []
        job = self.make_job(hcorn, hnorm3, keep)
        job.out_sharded = 1
        call = self.ctx.preprocess_host if self.world == 1 else self.ctx.preprocess_dist
        call(job)                                   # warm-up: grows the arenas
        self.ctx.L.cx_job_free(C.byref(job))
        barrier()
        t0 = time.perf_counter()
        br = {"weld_s": 0.0, "h2d_s": 0.0, "gpu_s": 0.0, "d2h_s": 0.0}
        nfl = 0
        for _ in range(steps):
            call(job)
            br["weld_s"] += job.t_weld_s; br["h2d_s"] += job.t_h2d_s; br["gpu_s"] += job.t_gpu_s; br["d2h_s"] += job.t_d2h_s
            assert int(job.nvert) == nv, (int(job.nvert), nv)
            nfl = int(job.nfluid)
            d2h_local = 16 * (job.vert_hi - job.vert_lo) + (16 * 3 + 4) * (job.seg_hi - job.seg_lo) + 4 * (job.vert_hi - job.vert_lo) + \
                4 * (job.fword_hi - job.fword_lo) + (4 * (self.ncell3 + 1) if self.rank == 0 else 0)
            self.ctx.L.cx_job_free(C.byref(job))
        barrier()
        dt = time.perf_counter() - t0
        fs_bytes = self.w["fshape"][0].shape[0] * 48 if self.w["fshape"] is not None else 0
        per = (nbe + self.world - 1) // self.world
        h2d_local = 48 * max(0, min(nbe, per * (self.rank + 1)) - min(nbe, per * self.rank)) + fs_bytes
        tt = torch.tensor([dt, float(h2d_local), float(d2h_local)], dtype=torch.float64, device=self.ctx.device)
        if self.world > 1:
            import torch.distributed as dist
            tmax = tt.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            dt = float(tmax[0].item())
        h2d, d2h = int(tt[1].item()), int(tt[2].item())
        api_name = ("cx_preprocess_host (raw STL corners + normals in pinned host memory; bounding box + weld on the device)" if self.world == 1 else
                    "cx_preprocess_dist on every rank (each uploads 1/N of the raw STL content, NCCL all-gather, device weld replicated, "
                    "calc_vert_volume and the fill sharded, each rank downloads its 1/N share of every result array)")
        return {"value": nbe * steps / dt, "unit": "triangles/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * dt / steps, "api": api_name, "breakdown_ms_rank0": {k: 1e3 * v / steps for k, v in br.items()}, "nfluid": nfl}
