#!/usr/bin/env python
"""The smallest program that runs the hot path device-resident, for ncu / compute-sanitizer:

    python tools/profile_pass.py CONFIG [SIZE] [PASSES]      e.g.  ncu ... python tools/profile_pass.py 2

CONFIG / SIZE as in bench.py --config / --size.  Prints the checksums of the last pass (so a profiled run can be compared with
a plain one)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from crixus_b200 import api, workloads as wl  # noqa: E402
from crixus_b200.dist import ShardedHotPath  # noqa: E402


def main():
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    size = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "-" else None
    passes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    ctx = api.Context(0)
    name, opts, gen = wl.bench_config(cfg, size)
    corners, normals = gen()
    nbe = corners.shape[0]
    dcorn = torch.from_numpy(corners.reshape(-1, 3)).cuda()
    pos4, ep4, lo, hi = ctx.weld_device(dcorn, opts["dr"])
    norm4 = torch.zeros((nbe, 4), dtype=torch.float32, device="cuda")
    norm4[:, :3] = torch.from_numpy(normals).cuda()
    w = dict(opts, name=name, nbe=nbe, nvert=int(pos4.shape[0]), pos=pos4.clone(), ep=ep4, norm=norm4, bbmin=lo, bbmax=hi)
    hp = ShardedHotPath(ctx, w, 0, 1, byte_stats=False)      # one pass inside
    for _ in range(passes):
        hp.step()
    torch.cuda.synchronize()
    print(name, "nbe", nbe, "nvert", w["nvert"], "alive", hp.nalive, "lattice", [int(x) for x in hp.dimg], hp.check(), flush=True)


if __name__ == "__main__":
    main()
