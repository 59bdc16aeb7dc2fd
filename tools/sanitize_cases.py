#!/usr/bin/env python
"""The program compute-sanitizer runs (one tool per gpurun call, B200_PROFILING.md):

    compute-sanitizer --tool memcheck  --log-file gpurun_out/memcheck.log  python tools/sanitize_cases.py
    compute-sanitizer --tool initcheck --log-file gpurun_out/initcheck.log python tools/sanitize_cases.py

smoke() (KAT plate + cube stage by stage + cx_preprocess_host) and three parity cases that cover the branches the BASELINE
workloads take -- spheric2_mini (container, box + geometry fill, free-surface triangles), channel (periodic links, periodic
neighbourhoods), sphere_tank (curved mesh) -- each compared with the oracle, plus the two-rank z-slab fill through the in-process
group (halo exchange, merge kernels)."""
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    import __graft_entry__ as ge
    ge.smoke()
    import cases
    from backends import CudaBackend
    from crixus_b200 import api
    from oracle import orc
    from test_gpu_parity import _cmp_pipeline
    G, O = CudaBackend(0), orc.Oracle()
    for name in ("spheric2_mini", "channel", "sphere_tank"):
        c = cases.case(name)
        o, g = cases.run(O, c), cases.run(G, c)
        _cmp_pipeline(o, g, c)
        print("case", name, "ok", flush=True)
    # two in-process ranks, z-slab form
    c = cases.case("cube_jitter")
    o = cases.run(O, c)
    dev = torch.device("cuda", 0)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    D = {k: d(o[k]) for k in ("fnorm", "fep", "fposa", "cell_idx")}
    pf = api.CxParams.from_buffer_copy(bytes(o["params_fill"]))
    f = [x for x in c["fills"] if x[0] == "geometry"][0]
    pf.dr_wall = np.float32(f[2])
    seed = api.seed_index(pf, f[1])
    ctxs = [api.Context(0, use_torch_stream=False) for _ in range(2)]
    api.dist_init_local(ctxs)
    out = [None, None]
    torch.cuda.synchronize()

    def work(r):
        ctxs[r].dist_set_fill_mode(2)
        fp = torch.zeros(o["fpos"].shape[0], dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        ctxs[r].dist_fill_geometry(pf, fp, D["fnorm"], D["fep"], D["fposa"], o["fnbe"], o["cnbe"], D["cell_idx"], seed)
        ctxs[r].synchronize()
        out[r] = fp.cpu().numpy().view(np.uint32)
    th = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    for r in range(2):
        assert np.array_equal(out[r], o["fpos"]), "rank %d" % r
    print("two-rank slab fill ok", flush=True)


if __name__ == "__main__":
    main()
