"""The multi-GPU path of the C-ABI (cx_dist_*, cx_preprocess_dist) against the single-GPU path, bit for bit.

Two ways to get N ranks:
  * in-process group (cx_dist_init_local): N contexts on ONE GPU, one host thread per rank.  Every line of the sharded algorithms
    (dealing of the directed tests, reduce to the slab owners, slab floods, halo exchange, termination, all-gather; interleaved
    vertex chunks; sharded upload / download of cx_preprocess_dist) runs here -- only the collective back-end differs from the
    real thing;
  * real NCCL ranks: two processes on two GPUs (skipped when fewer than two are visible), the library's own communicator
    (cx_dist_unique_id / cx_dist_init), compared with the single-GPU result of the same inputs.
"""
import ctypes as C
import os
import subprocess
import sys
import threading

import numpy as np
import pytest
import torch

import cases
from crixus_b200 import api
from test_gpu_pipeline import _arr, _job_for

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _run_ranks(n, fn):
    """fn(rank, ctx) on n threads, each with a context (own stream) of an in-process group; returns the per-rank results"""
    ctxs = [api.Context(0, use_torch_stream=False) for _ in range(n)]
    api.dist_init_local(ctxs)
    out, err = [None] * n, [None] * n

    def work(r):
        try:
            torch.cuda.set_device(0)
            out[r] = fn(r, ctxs[r])
            ctxs[r].synchronize()
        except BaseException as e:   # noqa: BLE001 -- reported below; a dead rank would otherwise hang the others silently
            err[r] = e
    th = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    assert not any(t.is_alive() for t in th), "a rank hangs"
    for e in err:
        if e is not None:
            raise e
    for c in ctxs:
        c.dist_finalize()
        c.close()
    return out


def _stage_inputs(o, dev):
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    return dict(pos=d(o["pos"]), norm=d(o["norm"]), ep=d(o["ep"]), trisize=d(o["trisize"]), cell_idx=d(o["cell_idx"]),
                fposa=d(o["fposa"]), fnorm=d(o["fnorm"]), fep=d(o["fep"]))


@pytest.mark.parametrize("name,n,mode", [("cube_jitter", 3, 2), ("spheric2_mini", 2, 2), ("spheric2_mini", 3, 1), ("sphere_tank", 4, 2),
                                         ("channel", 3, 2), ("cube_h05", 2, 0)])
def test_dist_stages_equal_single_gpu(name, n, mode, oracle):
    """cx_dist_calc_vert_volume and cx_dist_fill_geometry (z-slab form = mode 2, replicated form = mode 1) on n ranks: every rank
    ends with exactly the single-rank arrays, which are the oracle's."""
    c = cases.case(name)
    o = cases.run(oracle, c)
    nv, nbe = o["nvert"], o["nbe"]
    dev = torch.device("cuda", 0)
    D = _stage_inputs(o, dev)
    torch.cuda.synchronize()
    pv = api.CxParams.from_buffer_copy(bytes(o["params_volume"]))
    pf = api.CxParams.from_buffer_copy(bytes(o["params_fill"]))

    def rank_fn(r, ctx):
        ctx.dist_set_fill_mode(mode)
        vol = torch.full((nv,), 7.0, dtype=torch.float32, device=dev)     # garbage in: the call must overwrite everything
        torch.cuda.synchronize()
        ctx.dist_calc_vert_volume(pv, D["pos"], D["norm"], D["ep"], vol, D["trisize"], D["cell_idx"], nv, nbe, c["zero_open"])
        fpos = torch.zeros(o["fpos"].shape[0], dtype=torch.int32, device=dev)
        torch.cuda.synchronize()           # the context launches on a stream of its own
        total = 0
        for f in c["fills"]:
            if f[0] == "box":
                total += ctx.fill_box(pf, fpos, f[1], f[2])
                continue
            p = pf.copy()
            p.dr_wall = np.float32(f[2] if f[2] is not None else c["dr"])
            total += ctx.dist_fill_geometry(p, fpos, D["fnorm"], D["fep"], D["fposa"], o["fnbe"], o["cnbe"], D["cell_idx"], api.seed_index(p, f[1]))
        ctx.synchronize()
        return vol.cpu().numpy(), fpos.cpu().numpy().view(np.uint32), total

    for r, (vol, fpos, total) in enumerate(_run_ranks(n, rank_fn)):
        alive = o["pos"][:nv, 0] > -1e9
        np.testing.assert_array_equal(vol[alive].view(np.uint32), o["vol"][alive].view(np.uint32), err_msg="volumes, rank %d" % r)
        np.testing.assert_array_equal(fpos, o["fpos"], err_msg="fluid bits, rank %d" % r)
        assert total == o["nfluid"], (r, total, o["nfluid"])


@pytest.mark.parametrize("name,n,sharded", [("spheric2_mini", 3, 1), ("cube_jitter", 2, 0), ("channel", 2, 1), ("sphere_tank", 3, 1)])
def test_preprocess_dist_equals_preprocess_host(name, n, sharded, cuda_backend):
    """cx_preprocess_dist on n ranks (sharded upload, all-gather, sharded stages, sharded or gathered download) against
    cx_preprocess_host on one: the union of the ranks' shares is the single-GPU result, byte for byte."""
    c = cases.case(name)
    job0, keep0 = _job_for(c)
    cuda_backend.ctx.preprocess_host(job0)
    nv, nbe = int(job0.nvert), int(job0.nbe)
    G = int(job0.dimg[0] * job0.dimg[1] * job0.dimg[2])
    nw = (G + 31) // 32
    ref = dict(pos=_arr(job0.pos, (nv + nbe, 4), np.float32), norm=_arr(job0.norm, (nbe, 4), np.float32), ep=_arr(job0.ep, (nbe, 4), np.int32),
               surf=_arr(job0.surf, (nbe,), np.float32), vol=_arr(job0.vol, (nv,), np.float32), fpos=_arr(job0.fpos, (nw,), np.uint32),
               nfluid=int(job0.nfluid))
    cuda_backend.ctx.L.cx_job_free(C.byref(job0))

    def rank_fn(r, ctx):
        ctx.dist_set_fill_mode(2)
        got = None
        for rep in range(2):                   # second call reuses the arenas
            job, keep = _job_for(c)
            job.out_sharded = sharded
            ctx.preprocess_dist(job)
            assert int(job.nvert) == nv and int(job.nfluid) == ref["nfluid"]
            s0, s1, v0, v1, w0, w1 = (int(x) for x in (job.seg_lo, job.seg_hi, job.vert_lo, job.vert_hi, job.fword_lo, job.fword_hi))
            got = dict(seg=(s0, s1), vert=(v0, v1), fw=(w0, w1))
            if s1 > s0:
                got.update(bary=_arr(job.pos, (nv + nbe, 4), np.float32)[nv + s0:nv + s1], norm=_arr(job.norm, (nbe, 4), np.float32)[s0:s1],
                           ep=_arr(job.ep, (nbe, 4), np.int32)[s0:s1], surf=_arr(job.surf, (nbe,), np.float32)[s0:s1])
            if v1 > v0:
                got.update(vpos=_arr(job.pos, (nv + nbe, 4), np.float32)[v0:v1], vol=_arr(job.vol, (nv,), np.float32)[v0:v1])
            if w1 > w0:
                got.update(fpos=_arr(job.fpos, (nw,), np.uint32)[w0:w1])
            ctx.L.cx_job_free(C.byref(job))
        return got

    res = _run_ranks(n, rank_fn)
    cover = {k: np.zeros(m, dtype=np.int32) for k, m in (("seg", nbe), ("vert", nv), ("fw", nw))}
    for r, g in enumerate(res):
        for k in cover:
            cover[k][g[k][0]:g[k][1]] += 1
        s0, s1 = g["seg"]
        if s1 > s0:
            np.testing.assert_array_equal(g["bary"].view(np.uint32), ref["pos"][nv + s0:nv + s1].view(np.uint32))
            np.testing.assert_array_equal(g["norm"].view(np.uint32), ref["norm"][s0:s1].view(np.uint32))
            np.testing.assert_array_equal(g["ep"], ref["ep"][s0:s1])
            np.testing.assert_array_equal(g["surf"].view(np.uint32), ref["surf"][s0:s1].view(np.uint32))
        v0, v1 = g["vert"]
        if v1 > v0:
            np.testing.assert_array_equal(g["vpos"].view(np.uint32), ref["pos"][v0:v1].view(np.uint32))
            np.testing.assert_array_equal(g["vol"].view(np.uint32), ref["vol"][v0:v1].view(np.uint32))
        w0, w1 = g["fw"]
        if w1 > w0:
            np.testing.assert_array_equal(g["fpos"], ref["fpos"][w0:w1])
    for k, cnt in cover.items():          # every element downloaded exactly once (sharded) / by rank 0 alone (gathered)
        assert (cnt == 1).all(), k


_NCCL_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from crixus_b200 import api, workloads as wl
from crixus_b200.dist import ShardedHotPath
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))   # ships the 128-byte id and reduces the timings; the data path is the library's own communicator
ctx = api.Context(rank)
box = [api.dist_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
ctx.dist_init(box[0], world, rank)
out = {}
for cfg, size, mode in ((4, 120, 2), (2, 1, 1), (3, 96, 2)):
    name, opts, gen = wl.bench_config(cfg, size)
    corners, normals = gen()
    nbe = corners.shape[0]
    dcorn = torch.from_numpy(corners.reshape(-1, 3)).cuda()
    pos4, ep4, lo, hi = ctx.weld_device(dcorn, opts["dr"])
    norm4 = torch.zeros((nbe, 4), dtype=torch.float32, device="cuda"); norm4[:, :3] = torch.from_numpy(normals).cuda()
    w = dict(opts, name=name, nbe=nbe, nvert=int(pos4.shape[0]), pos=pos4.clone(), ep=ep4, norm=norm4, bbmin=lo, bbmax=hi)
    ctx.dist_set_fill_mode(mode)
    hp = ShardedHotPath(ctx, w, rank, world, byte_stats=False)
    hp.step()
    torch.cuda.synchronize()
    chk = hp.check()
    ref = hp.single_gpu_check()                        # the same inputs through the single-GPU calls, on every rank
    assert all(ref[k] == chk[k] for k in ref), (cfg, rank, chk, ref)
    hcorn = torch.from_numpy(corners.reshape(-1, 3)).pin_memory(); hnorm = torch.from_numpy(normals).pin_memory()
    e = hp.e2e(1, lambda: (dist.barrier(), torch.cuda.synchronize()), hcorn, hnorm)
    assert e["nfluid"] == chk["nfluid_popcount"], (e["nfluid"], chk)
    out[cfg] = chk
print("RANK%d OK %s" % (rank, out), flush=True)
ctx.dist_finalize()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (real NCCL ranks); the in-process tests above cover one GPU")
def test_two_real_nccl_ranks_equal_single_gpu(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_NCCL_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29641",
           str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "RANK0 OK" in r.stdout and "RANK1 OK" in r.stdout
