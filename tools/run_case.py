"""Runs one parity case (tests/cases.py) on the CUDA path and prints a stage-by-stage comparison with the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from backends import CudaBackend  # noqa: E402
from oracle import orc  # noqa: E402

names = sys.argv[1:] or cases.ALL
G, O = CudaBackend(0), orc.Oracle()
bad = 0
for name in names:
    c = cases.case(name)
    o = cases.run(O, c)
    g = cases.run(G, c)
    nv = o["nvert"]
    alive = o["pos"][:nv, 0] > -1e9
    q = float(c["dr"] / 40) ** 3
    rep = {}
    for k in ("zstat", "pre_ep", "pre_surf", "pre_bary", "cell_idx", "ep", "pos", "norm", "surf", "trisize", "fpos"):
        if k in o:
            rep[k] = bool(np.array_equal(np.asarray(g[k])[..., :3] if np.asarray(g[k]).ndim == 2 else g[k],
                                         np.asarray(o[k])[..., :3] if np.asarray(o[k]).ndim == 2 else o[k]))
    dv = (g["vol"] != o["vol"]) & alive
    rep["vol_diff"] = int(dv.sum())
    rep["vol_maxq"] = float(np.abs(g["vol"] - o["vol"])[alive].max() / q)
    rep["nfluid"] = (g.get("nfluid"), o.get("nfluid"), g.get("fill_counts"), o.get("fill_counts"))
    ok = all(v for k, v in rep.items() if isinstance(v, bool)) and rep["vol_diff"] == 0 and g.get("nfluid") == o.get("nfluid")
    bad += not ok
    print(name, "OK" if ok else "MISMATCH", rep, "rounds", G.last_sweeps, flush=True)
    if dv.any():
        ii = np.nonzero(dv)[0][:8]
        print("   first vol diffs (idx, gpu, oracle, trisize):", [(int(i), float(g["vol"][i]), float(o["vol"][i]), int(o["trisize"][i])) for i in ii])
sys.exit(1 if bad else 0)
