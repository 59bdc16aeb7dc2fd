"""Generates the golden fixtures tests/golden/*.npz by running the UNMODIFIED reference CUDA program
(oracle/_ref/Crixus_ref, built by oracle/build_ref.sh from /root/reference) on a B200:

    gpurun -- python tests/golden/make_golden.py gpurun_out/golden      # then copy gpurun_out/golden/*.npz here

For every case of tests/cases.py it writes the binary STL(s) + the .ini the reference reads, runs
`Crixus_ref case.ini` ([output] format=vtu), parses the .vtu and stores what parity is judged on:
vertex volumes/positions (vertex numbering is deterministic), the segment records in canonical order
(the reference's segment order is run-dependent, crixus_d.cu:94), and the fluid particle positions.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from crixus_b200 import meshgen as mg  # noqa: E402
from vtu import read_vtu, split_particles  # noqa: E402


def write_ini(path, c, stl, fshape_path, sb_paths=(), split=False):
    L = ["[mesh]", "stlfile=%s" % stl, "dr=%.9g" % float(c["dr"]), "swap_normals=%s" % ("true" if c["swap_normals"] else "false"),
         "zeroOpen=%s" % ("true" if c["zero_open"] else "false")]
    if fshape_path:
        L.append("fshape=%s" % fshape_path)
    L += ["[periodicity]"] + ["%s=%s" % (a, "true" if c["periodic"][i] else "false") for i, a in enumerate("xyz")]
    if c["container"] is not None:
        lo, hi = c["container"]
        L += ["[fluid_container]", "use=true"] + ["%smin=%.9g" % (a, lo[i]) for i, a in enumerate("xyz")] + \
             ["%smax=%.9g" % (a, hi[i]) for i, a in enumerate("xyz")]
    for k, f in enumerate(c["fills"]):
        L.append("[fill_%d]" % k)
        if f[0] == "box":
            L += ["option=box"] + ["%smin=%.9g" % (a, f[1][i]) for i, a in enumerate("xyz")] + \
                 ["%smax=%.9g" % (a, f[2][i]) for i, a in enumerate("xyz")]
        else:
            L += ["option=geometry"] + ["%sseed=%.9g" % (a, f[1][i]) for i, a in enumerate("xyz")] + ["dr_wall=%.9g" % float(f[2])]
    if sb_paths:
        L += ["[special_boundary_grids]"] + ["mesh%d=%s" % (k + 1, sp) for k, sp in enumerate(sb_paths)]
    L += ["[output]", "format=vtu", "name=%s" % c["name"]] + (["split=true"] if split else [])
    open(path, "w").write("\n".join(L) + "\n")


def write_special(work, c):
    """the special-boundary grids of a case as binary STL files (facet normals are not read, crixus.cu:538-556)"""
    paths = []
    for k, sb in enumerate(c.get("special", ())):
        sp = os.path.join(work, "%s_sbgrid_%d.stl" % (c["name"], k + 1))
        mg.write_stl(sp, sb, np.zeros((sb.shape[0], 3), dtype=np.float32))
        paths.append(sp)
    return paths


def canonical_segments(seg):
    key = np.lexsort((seg["Points"][:, 2], seg["Points"][:, 1], seg["Points"][:, 0], seg["VertexParticle"][:, 2],
                      seg["VertexParticle"][:, 1], seg["VertexParticle"][:, 0]))
    return {k: v[key] for k, v in seg.items()}


def main(outdir):
    outdir = os.path.abspath(outdir)
    work = os.path.join(outdir, "work")
    os.makedirs(work, exist_ok=True)
    exe = os.path.join(ROOT, "oracle", "_ref", "Crixus_ref")
    names = sys.argv[2:] or cases.ALL
    for name in names:
        c = cases.case(name)
        stl = os.path.join(work, name + ".stl")
        mg.write_stl(stl, c["tris"], c["normals"])
        fsp = None
        if c["fshape"] is not None:
            fsp = os.path.join(work, name + "_fs.stl")
            mg.write_stl(fsp, c["fshape"][0], c["fshape"][1])
        sbp = write_special(work, c)
        ini = os.path.join(work, name + ".ini")
        write_ini(ini, c, stl, fsp, sbp)
        log = subprocess.run([exe, ini], cwd=work, capture_output=True, text=True, timeout=1200)
        open(os.path.join(work, name + ".log"), "w").write(log.stdout + log.stderr)
        d = split_particles(read_vtu(os.path.join(work, name + ".vtu")))
        seg = canonical_segments(d["segment"])
        np.savez_compressed(os.path.join(outdir, name + ".npz"),
                            vert_pos=d["vertex"]["Points"].astype(np.float32), vert_vol=d["vertex"]["Volume"],
                            seg_pos=seg["Points"].astype(np.float32), seg_norm=seg["Normal"], seg_surf=seg["Surface"],
                            seg_ep=seg["VertexParticle"].astype(np.int64), fluid_pos=d["fluid"]["Points"].astype(np.float32),
                            vert_kent=d["vertex"]["KENT"].astype(np.int32), seg_kent=seg["KENT"].astype(np.int32),
                            seg_kent_file_order=d["segment"]["KENT"].astype(np.int32),
                            fluid_vol=d["fluid"]["Volume"][:1], n_file=np.int64(read_vtu(os.path.join(work, name + ".vtu"))["n"]),
                            log=np.array(log.stdout))
        print(name, "verts", d["vertex"]["Points"].shape[0], "segs", seg["Points"].shape[0], "fluid", d["fluid"]["Points"].shape[0], flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
