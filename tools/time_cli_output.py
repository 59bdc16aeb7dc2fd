#!/usr/bin/env python
"""Times our command line on BASELINE configs[1] (3.06 M triangles, 11.6 M particles, 1.1 GB .vtu) with the device-side
record assembly + streamed writers (default) and with the host restatement (CX_HOST_ASSEMBLY=1); checks that both write the
same bytes.      gpurun -- python tools/time_cli_output.py gpurun_out/cli_output_timing.json [level]"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from crixus_b200 import meshgen as mg  # noqa: E402
from make_golden import write_ini  # noqa: E402
from time_reference_cuda import case_for  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "cli_output_timing.json")
    level = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    c = case_for(level)
    c["name"] = "m"
    exe = os.path.join(ROOT, "crixus_b200", "bin", "crixus")
    res = {"level": level, "triangles": int(c["tris"].shape[0])}
    with tempfile.TemporaryDirectory() as d:
        stl, fsp, ini = os.path.join(d, "m.stl"), os.path.join(d, "m_fs.stl"), os.path.join(d, "m.ini")
        mg.write_stl(stl, c["tris"], c["normals"])
        mg.write_stl(fsp, c["fshape"][0], c["fshape"][1])
        for fmt in ("vtu", "h5sph"):
            write_ini(ini, c, stl, fsp)
            text = open(ini).read().replace("format=vtu", "format=" + fmt)
            open(ini, "w").write(text)
            fn = os.path.join(d, "m.vtu" if fmt == "vtu" else "0.m.h5sph")
            for mode in ("device", "host", "device", "host"):
                env = dict(os.environ, CX_HOST_ASSEMBLY="1" if mode == "host" else "0")
                t0 = time.perf_counter()
                r = subprocess.run([exe, ini], cwd=d, capture_output=True, text=True, env=env, timeout=600)
                wall = time.perf_counter() - t0
                assert r.returncode == 0, r.stdout[-2000:]
                h = hashlib.sha256(open(fn, "rb").read()).hexdigest()
                row = {"wall_s": wall, "bytes": os.path.getsize(fn), "sha256": h,
                       "log": [ln.strip() for ln in r.stdout.splitlines() if "Timing:" in ln]}
                res.setdefault(fmt, {}).setdefault(mode, []).append(row)
                print(fmt, mode, "%.3f s" % wall, row["log"][-1], flush=True)
                os.remove(fn)
            hs = {row["sha256"] for m in res[fmt].values() for row in m}
            res[fmt]["identical"] = len(hs) == 1
            assert len(hs) == 1, "device and host paths wrote different files"
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
