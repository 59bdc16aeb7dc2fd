import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from crixus_b200 import api, meshgen as mg, workloads as wl
from crixus_b200.dist import ShardedHotPath
ctx = api.Context(0)
w = wl.spheric2_family(wl.SPHERIC2_DR / 2, 1.25, levels=2, base_dr=wl.SPHERIC2_DR)
w["fills"] = []
w["name"] = "l2"
w = wl._finish(w)
hp = ShardedHotPath(ctx, w, 0, 1)
for _ in range(4): hp.step(timed=True)
torch.cuda.synchronize()
ms = hp.collect_stage_ms() / 4
print(sys.argv[1:], "vv %.2f ms" % ms[hp.STAGES.index("vert_volume")], flush=True)
