#!/usr/bin/env python
"""Times the UNMODIFIED reference CUDA program (oracle/_ref/Crixus_ref, built by oracle/build_ref.sh for sm_100) and our
command line on the same .ini + STL files, on the GPU box:

    gpurun -- python tools/time_reference_cuda.py gpurun_out/ref_timing.json [levels ...]

levels: midpoint subdivisions of the SPHERIC-2 mesh (0 = BASELINE configs[0] = spheric2.ini, 3 = configs[1]).
Wall-clock of the whole process (file in -> file out); the reference has no internal timers.  A reported baseline."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import numpy as np  # noqa: E402

from crixus_b200 import meshgen as mg, workloads as wl  # noqa: E402
from make_golden import write_ini  # noqa: E402


def case_for(level):
    w = wl.spheric2_family(wl.SPHERIC2_DR / 2 ** max(level - 1, 0) if level else wl.SPHERIC2_DR, 1.25, levels=level, base_dr=wl.SPHERIC2_DR)
    if level == 0:
        w["fills"] = [("box", (0.2, 0.2, 0.2), (0.4, 0.4, 0.4)), ("geometry", (0.5, 0.5, 0.5), np.float32(wl.SPHERIC2_DR))]
    w.update(name="spheric2_l%d" % level, periodic=(False, False, False), zero_open=False)
    return w


def run(cmd, cwd, limit):
    t0 = time.perf_counter()
    try:
        r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=limit)
        return time.perf_counter() - t0, r.returncode, r.stdout[-2000:]
    except subprocess.TimeoutExpired:
        return None, None, "timeout after %d s" % limit


def compare_outputs(d, c, ini):
    """runs both programs once more into separate directories and compares the particle files (segments in canonical
    order: the reference's intra-cell segment order is run-dependent, crixus_d.cu:94)"""
    import re
    import shutil
    from make_golden import canonical_segments
    from vtu import read_vtu, split_particles
    outs = {}
    for name, exe in (("ours", os.path.join(ROOT, "crixus_b200", "bin", "crixus")), ("ref", os.path.join(ROOT, "oracle", "_ref", "Crixus_ref"))):
        wd = os.path.join(d, name)
        os.makedirs(wd, exist_ok=True)
        r = subprocess.run([exe, ini], cwd=wd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, name + " failed"
        p = split_particles(read_vtu(os.path.join(wd, c["name"] + ".vtu")))
        p["segment"] = canonical_segments(p["segment"])
        p["counted"] = int(re.search(r"Creation of (\d+) fluid particles", r.stdout).group(1))
        outs[name] = p
        shutil.rmtree(wd)
    a, b = outs["ours"], outs["ref"]
    nf = b["fluid"]["Points"].shape[0]
    shift = b["counted"] - nf              # the reference's racy extra count: 0 or 1
    assert shift in (0, 1), "reference count shift %d" % shift
    assert a["counted"] == nf, "fluid count %d vs %d" % (a["counted"], nf)
    for grp, keys in (("fluid", ("Points", "Volume")), ("vertex", ("Points", "Volume")), ("segment", ("Points", "Normal", "Surface"))):
        for k in keys:
            assert np.array_equal(a[grp][k], b[grp][k]), "%s.%s differs" % (grp, k)
    assert np.array_equal(a["segment"]["VertexParticle"].astype(np.int64) + shift, b["segment"]["VertexParticle"].astype(np.int64)), "segment links differ"
    return {"bit_exact": True, "fluid": int(nf), "vertices": int(a["vertex"]["Points"].shape[0]), "segments": int(a["segment"]["Points"].shape[0]),
            "reference_count_shift": int(shift)}


def compare_level(d, level, parity=True, timing=True, limit=300):
    """one SPHERIC-2 level in directory d: both programs on the same .ini + STL files -> dict (wall clocks, parity)"""
    lv = level
    c = case_for(lv)
    c["name"] = "m"
    stl, fsp, ini = os.path.join(d, "m.stl"), os.path.join(d, "m_fs.stl"), os.path.join(d, "m.ini")
    mg.write_stl(stl, c["tris"], c["normals"])
    mg.write_stl(fsp, c["fshape"][0], c["fshape"][1])
    write_ini(ini, c, stl, fsp)
    row = {"level": lv, "triangles": int(c["tris"].shape[0]), "dr": float(c["dr"])}
    ok = True
    for name, exe in (("ours", os.path.join(ROOT, "crixus_b200", "bin", "crixus")), ("reference_cuda", os.path.join(ROOT, "oracle", "_ref", "Crixus_ref"))):
        if not os.path.exists(exe):
            row[name] = "missing"
            ok = False
            continue
        if not timing:
            continue
        run([exe, ini], d, limit) if name == "ours" else None      # warm-up of our side only (driver/context start-up is in both)
        t, rc, tail = run([exe, ini], d, limit)
        row[name + "_wall_s"] = t
        row[name + "_rc"] = rc
        ok = ok and rc == 0
        for line in tail.splitlines():
            if "Timing:" in line or "fluid particles completed" in line:
                row.setdefault(name + "_log", []).append(line.strip())
    # ---- full-size parity: our .vtu against the reference's, particle by particle
    if parity and ok:
        try:
            row["parity"] = compare_outputs(d, c, ini)
        except AssertionError as e:
            row["parity"] = "MISMATCH: " + str(e)[:300]
    return row


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_timing.json")
    levels = [int(x) for x in sys.argv[2:]] or [0, 1, 2]
    limit = int(os.environ.get("CX_REF_TIMEOUT", "300"))
    res = []
    for lv in levels:
        with tempfile.TemporaryDirectory() as d:
            row = compare_level(d, lv, limit=limit)
            res.append(row)
            print(json.dumps(row), flush=True)
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
