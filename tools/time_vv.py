#!/usr/bin/env python
"""calc_vert_volume alone on a BASELINE workload, device resident, CUDA-event timed (kernel experiments: CX_LIB_PATH picks the
library build):   python tools/time_vv.py CONFIG [SIZE] [REPS]"""
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
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    ctx = api.Context(0)
    name, opts, gen = wl.bench_config(cfg, size)
    corners, normals = gen()
    nbe = corners.shape[0]
    dcorn = torch.from_numpy(corners.reshape(-1, 3)).cuda()
    pos4, ep4, lo, hi = ctx.weld_device(dcorn, opts["dr"])
    norm4 = torch.zeros((nbe, 4), dtype=torch.float32, device="cuda")
    norm4[:, :3] = torch.from_numpy(normals).cuda()
    w = dict(opts, name=name, nbe=nbe, nvert=int(pos4.shape[0]), pos=pos4.clone(), ep=ep4, norm=norm4, bbmin=lo, bbmax=hi)
    hp = ShardedHotPath(ctx, w, 0, 1, byte_stats=False)
    hp.step()
    torch.cuda.synchronize()
    for _ in range(reps):
        hp.step(timed=True)
    torch.cuda.synchronize()
    ms = hp.collect_stage_ms() / reps
    vv = float(ms[hp.STAGES.index("vert_volume")])
    fill = float(ms[hp.STAGES.index("fill")])
    chk = hp.check()
    print("%s | lib %s | nvert %d | vert_volume %.3f ms  %.2f Mverts/s | vol_hash %s | fill %.3f ms fpos_hash %s" % (
        name, os.path.basename(os.environ.get("CX_LIB_PATH", "default")), w["nvert"], vv, w["nvert"] / vv / 1e3, chk["vol_hash"], fill,
        chk["fpos_hash"]), flush=True)


if __name__ == "__main__":
    main()
