"""Multi-rank host logic (crixus_b200/dist.py) on CPU: gloo, world_size 2 and 3, the CPU oracle as the compute
backend.  What is checked is the sharding itself: vertex ranges + all-gather reproduce the single-rank volumes, and the
z-slab fill with halo exchange reaches the same fixed point (bit-exact flags and counts) as the unsharded fill."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from crixus_b200 import dist as cd   # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cases
        from oracle import orc
        O = orc.Oracle()
        c = cases.case(name)
        ref = cases.run(O, c)                       # every rank: the unsharded oracle run (also provides the stage inputs)
        nvert, nbe = ref["nvert"], ref["nbe"]
        # ---- calc_vert_volume: contiguous vertex ranges balanced by trisize, volumes all-gathered
        sizes = [cd.split_weighted(ref["trisize"], world, r) for r in range(world)]
        v0, v1 = sizes[rank]
        part = O.calc_vert_volume(ref["params_volume"], ref["pos"], ref["norm"], ref["ep"], ref["trisize"], ref["cell_idx"], nvert, nbe,
                                  c["zero_open"], v0, v1)
        mx = max(b - a for a, b in sizes)
        buf = torch.zeros(mx, dtype=torch.float32)
        buf[: v1 - v0] = torch.from_numpy(part[v0:v1])
        out = [torch.zeros(mx, dtype=torch.float32) for _ in range(world)]
        dist.all_gather(out, buf)
        vol = np.concatenate([o.numpy()[: b - a] for o, (a, b) in zip(out, sizes)])
        assert sizes[0][0] == 0 and sizes[-1][1] == nvert and all(sizes[i][1] == sizes[i + 1][0] for i in range(world - 1))
        assert np.array_equal(vol.view(np.uint32), ref["vol"].view(np.uint32)), "sharded volumes differ"
        # ---- fills: box fills replicated, geometry fill in z-slabs with halo exchange
        p = ref["params_fill"].copy()
        dimg = ref["dimg"]
        fpos = torch.zeros(ref["fpos"].shape[0], dtype=torch.int32)
        fnp = fpos.numpy().view(np.uint32)
        total = 0
        for f in c["fills"]:
            if f[0] == "box":
                total += O.fill_box(p, fnp, [np.float32(x) for x in f[1]], [np.float32(x) for x in f[2]])
                continue
            p.dr_wall = np.float32(f[2] if f[2] is not None else c["dr"])
            seed = O.seed_index(p, f[1])

            def slab(fp, seed_, k0, k1, first):
                return O.fill_geometry_slab(p, fp.numpy().view(np.uint32), ref["fnorm"], ref["fep"], ref["fposa"], ref["fnbe"],
                                            ref["cnbe"], ref["cell_idx"], seed_, k0, k1, first)
            n, its = cd.distributed_fill(slab, fpos, dimg, seed, rank, world, dist)
            cd.gather_slabs(fpos, dimg, rank, world, dist)
            total += n
            assert its >= 1
        assert np.array_equal(fnp, ref["fpos"]), "sharded fill flags differ"
        assert total == ref["nfluid"], (total, ref["nfluid"])
        open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("cube_jitter", 2), ("spheric2_mini", 2), ("sphere_tank", 3)])
def test_sharded_equals_single_rank(name, world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / ("ok%d" % r)) for r in range(world))


def test_partition_helpers():
    for n, w in ((10, 3), (7, 8), (131, 4), (0, 2)):
        r = [cd.split_range(n, w, k) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    wts = np.array([1, 1, 50, 1, 1, 1, 1, 50, 1, 1])
    r = [cd.split_weighted(wts, 2, k) for k in range(2)]
    assert r[0][0] == 0 and r[1][1] == len(wts) and r[0][1] == r[1][0]
    assert abs(int(wts[r[0][0]:r[0][1]].sum()) - int(wts[r[1][0]:r[1][1]].sum())) <= 50
    # plane_words covers every bit of the plane
    dimg = (13, 7, 5)
    for k in range(5):
        a, b = cd.plane_words(dimg, k)
        assert a * 32 <= k * 91 and b * 32 >= (k + 1) * 91
    # own_bits_only keeps exactly the rank's bit range, also across unaligned word boundaries
    G = 13 * 7 * 5
    full = torch.full(((G + 31) // 32,), -1, dtype=torch.int32)
    acc = torch.zeros_like(full)
    for rank in range(3):
        k0, k1 = cd.split_range(5, 3, rank)
        part = cd.own_bits_only(full.clone(), dimg, k0, k1)
        bits = np.unpackbits(part.numpy().view(np.uint8), bitorder="little")
        want = np.zeros_like(bits)
        want[k0 * 91:k1 * 91] = 1
        assert np.array_equal(bits[:G], want[:G])
        acc += part
    assert np.array_equal(np.unpackbits(acc.numpy().view(np.uint8), bitorder="little")[:G], np.ones(G, dtype=np.uint8))


@pytest.mark.parametrize("nparts", [2, 3, 4, 8])
def test_dealing_of_the_directed_tests_is_a_balanced_partition(nparts):
    """The multi-GPU fill merges the blocked planes of the ranks with SUM reductions, which equal OR only if every lattice cell
    is resolved by exactly one rank; and the walls of a tank -- whole planes of the 3dr-cell grid -- must not all land on a few
    ranks.  cd.dealt_chunk_owner mirrors the kernel's formula (csrc/fill.cu chunk_owner: chunks of 32 consecutive 3dr-cells)."""
    gridn = (100, 100, 100)
    ncell = gridn[0] * gridn[1] * gridn[2]
    nchunks = (ncell + 31) // 32
    owner = cd.dealt_chunk_owner(np.arange(nchunks), nparts)
    # partition: within every complete group of nparts chunks each rank appears exactly once
    full = (nchunks // nparts) * nparts
    grp = owner[:full].reshape(-1, nparts)
    assert (np.sort(grp, axis=1) == np.arange(nparts)).all()
    # the kernel's enumeration (rank `part` takes slot (part + rot(g)) % nparts of group g) hits exactly the chunks it owns
    g = np.arange(nchunks // nparts)
    rot = ((g.astype(np.uint64) * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)) >> np.uint64(16)
    for part in range(nparts):
        mine = g * nparts + (part + (rot % np.uint64(nparts)).astype(np.int64)) % nparts
        assert (owner[mine] == part).all()
    # balance on the six walls of the tank (two layers of 3dr-cells each hold the lattice cells whose tests are expensive)
    A, B = np.meshgrid(np.arange(gridn[0]), np.arange(gridn[1]), indexing="ij")
    walls = []
    for layer in (0, 1, gridn[2] - 2, gridn[2] - 1):
        walls.append(cd.cell3_chunk(A.ravel(), B.ravel(), np.full(A.size, layer), gridn))
        walls.append(cd.cell3_chunk(A.ravel(), np.full(A.size, layer), B.ravel(), gridn))
        walls.append(cd.cell3_chunk(np.full(A.size, layer), A.ravel(), B.ravel(), gridn))
    load = np.bincount(cd.dealt_chunk_owner(np.concatenate(walls), nparts), minlength=nparts)
    assert load.max() <= 1.15 * load.mean(), load
    for one in walls[:3]:                                                 # every single wall layer by itself
        load = np.bincount(cd.dealt_chunk_owner(one, nparts), minlength=nparts)
        assert load.max() <= 1.3 * load.mean(), load
