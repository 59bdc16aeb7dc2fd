#!/usr/bin/env python
"""calc_vert_volume rate on meshes that stress different regimes of the kernel (gpurun -- python tools/bench_vert_volume_meshes.py):
closed cube tanks with h/dr = 1.25, 0.625, 0.5 (commensurate with the voxel grid: planes on voxel planes and through the
faces of the sampling cube), 0.53 (generic), a shifted copy, a sphere in a tank, and the SPHERIC-2 tank of the bench."""
import sys, time, numpy as np, torch
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
    ts = hp.trisize.cpu().numpy()
    print("%-28s tris %8d verts %8d  T mean %.2f max %d  vv %.2f ms  %.2f Mverts/s" % (name, w["nbe"], w["nvert"], ts.mean(), ts.max(), vv, w["nvert"] / vv / 1e3), flush=True)
for hf in (1.25, 0.625, 0.5, 0.53):
    dr = 1.0 / 150
    v, t, n = mg.box_tank((0, 0, 0), (1, 1, 1), hf * dr, inward=False)
    run("cube tank hfac=%.3f" % hf, v, t, n, dr)
dr = 1.0 / 150
v, t, n = mg.box_tank((0, 0, 0), (1, 1, 1), 0.5 * dr, inward=False)
v2 = v + np.float32(0.013)
run("cube tank 0.5, shifted", v2, t, n, dr)
v, t, n = mg.sphere_in_tank(side_cells=150, dr=dr, hfac=0.5, sphere_levels=5)
run("sphere tank 0.5", v, t, n, dr)
w = wl.spheric2_family(wl.SPHERIC2_DR / 2, 1.25, levels=2, base_dr=wl.SPHERIC2_DR)
run("spheric2 l2", w["tris"].reshape(-1,3), np.arange(w["tris"].shape[0]*3).reshape(-1,3), w["normals"], w["dr"])
