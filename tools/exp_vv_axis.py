import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from crixus_b200 import api, meshgen as mg, workloads as wl
from crixus_b200.dist import ShardedHotPath
ctx = api.Context(0)
def run(name, v, t, n, dr, swap=True):
    w = dict(tris=v[t], normals=n, dr=np.float32(dr), swap_normals=swap, fills=[], name=name)
    w = wl._finish(w)
    hp = ShardedHotPath(ctx, w, 0, 1)
    hp.step(timed=True); hp.step(timed=True)
    torch.cuda.synchronize()
    ms = hp.collect_stage_ms() / 2
    vv = ms[hp.STAGES.index("vert_volume")]
    print("%-28s tris %8d verts %8d  vv %.2f ms  %.2f Mverts/s" % (name, w["nbe"], w["nvert"], vv, w["nvert"] / vv / 1e3), flush=True)
dr = 1.0 / 300
for hf in (0.625, 0.5):
    for name, hi in (("flat z", (1, 1, 0.04)), ("flat y", (1, 0.04, 1)), ("flat x", (0.04, 1, 1))):
        v, t, n = mg.box_tank((0, 0, 0), hi, hf * dr, inward=False)
        run("%s hfac=%.3f" % (name, hf), v, t, n, dr)
