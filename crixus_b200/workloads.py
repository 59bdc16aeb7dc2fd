"""The synthetic workloads of BASELINE.json `configs`, as welded meshes ready for upload.

A workload is a dict with the STL content (tris, normals), the options of the .ini that would drive the reference,
and the welded mesh (pos float4, ep int4) exactly as crixus_main holds it before its first cudaMemcpy
(/root/reference/src/crixus.cu:225-245)."""
from __future__ import annotations

import numpy as np

from . import api, meshgen as mg

SPHERIC2_DR = 0.018333   # resources/spheric2.ini


def _finish(w):
    tris = np.ascontiguousarray(w["tris"], dtype=np.float32)
    corners = tris.reshape(-1, 3)
    w["bbmin"], w["bbmax"] = corners.min(0), corners.max(0)
    verts, epv = api.weld_host(corners, np.float32(w["dr"]))
    nv, nbe = verts.shape[0], tris.shape[0]
    pos = np.zeros((nv, 4), dtype=np.float32)
    pos[:, :3] = verts
    ep = np.zeros((nbe, 4), dtype=np.int32)
    ep[:, :3] = epv
    norm = np.zeros((nbe, 4), dtype=np.float32)
    norm[:, :3] = w["normals"]
    w.update(tris=tris, nvert=nv, nbe=nbe, pos=pos, ep=ep, norm=norm)
    w.setdefault("periodic", (False, False, False))
    w.setdefault("zero_open", False)
    w.setdefault("container", None)
    w.setdefault("fshape", None)
    return w


def spheric2_family(dr, hfac, levels=0, base_dr=None):
    """SPHERIC-2 tank with the shipped .ini options: swap_normals, container (0,0,0)-(1.5,1,0.6), geometry fill seeded
    at (0.5,0.5,0.5) with dr_wall = dr and the shipped 4-triangle free-surface shape."""
    v, t, n = mg.spheric2(dr=base_dr if base_dr is not None else dr, hfac=hfac)
    if levels:
        v, t, n = mg.subdivide(v, t, n, levels)
    fv, ft, fn = mg.spheric2_fshape()
    return dict(tris=v[t], normals=n, dr=np.float32(dr), swap_normals=True, container=((0.0, 0.0, 0.0), (1.5, 1.0, 0.6)),
                fills=[("geometry", (0.5, 0.5, 0.5), np.float32(dr))], fshape=(fv[ft], fn))


def config1():
    """configs[0]: resources/spheric2.ini verbatim (box fill + geometry fill), ~4.8e4 triangles."""
    w = spheric2_family(SPHERIC2_DR, 1.25)
    w["fills"] = [("box", (0.2, 0.2, 0.2), (0.4, 0.4, 0.4)), ("geometry", (0.5, 0.5, 0.5), np.float32(SPHERIC2_DR))]
    w["name"] = "config1: spheric2.ini, %d triangles" % w["tris"].shape[0]
    return _finish(w)


def config2(levels=3):
    """configs[1]: spheric2 midpoint-subdivided x64 (~3.1e6 triangles) at dr/4, fluid filled via the fshape mode."""
    w = spheric2_family(SPHERIC2_DR / 2 ** (levels - 1), 1.25, levels=levels, base_dr=SPHERIC2_DR)
    w["name"] = "config2: spheric2 subdivided x%d, dr=%.6g, fshape geometry fill" % (4 ** levels, float(w["dr"]))
    return _finish(w)


def config2_sample(dr):
    """A bounded sample of config 2 for the CPU baseline: same geometry, options and mesh refinement relative to dr
    (h = 0.625 dr) at a coarser dr, so that the CPU oracle finishes in seconds."""
    w = spheric2_family(dr, 0.625)
    w["name"] = "config2 sample: spheric2 at dr=%.6g, h=0.625dr, %d triangles" % (dr, w["tris"].shape[0])
    return _finish(w)


def channel(n=1581, dr=0.01):
    """configs[2]: x/y-periodic channel, 2 planes of n x n quads (~1e7 triangles at n=1581)."""
    v, t, nn = mg.channel(n=n, dr=dr)
    L = n * 1.25 * dr
    H = float(v[:, 2].max())
    w = dict(tris=v[t], normals=nn, dr=np.float32(dr), swap_normals=True, periodic=(True, True, False),
             fills=[("geometry", (L / 2, L / 2, H / 2), np.float32(dr))])
    w["name"] = "config3: periodic channel n=%d" % n
    return _finish(w)


def sphere_tank(side_cells=200, dr=0.005, sphere_levels=6):
    """configs[3] family: closed cubic tank with a sphere; side_cells^3 lattice cells."""
    v, t, n = mg.sphere_in_tank(side_cells=side_cells, dr=dr, hfac=0.5, sphere_levels=sphere_levels)
    w = dict(tris=v[t], normals=n, dr=np.float32(dr), swap_normals=True,
             fills=[("geometry", (4 * dr, 4 * dr, 4 * dr), np.float32(dr))])
    w["name"] = "config4 family: sphere in tank, %d^3 cells" % side_cells
    return _finish(w)


# ------------------------------------------------------------------------------------------------ BASELINE.json configs
# bench.py --config N (1-based, N = index into BASELINE.json `configs` + 1).  Small workloads are welded on the host
# (_finish); the large ones are only GENERATED here (raw STL content) and welded on the device by the harness.
def _options(dr, fills, periodic=(False, False, False), container=None, fshape=None, swap_normals=True, zero_open=False):
    return dict(dr=np.float32(dr), swap_normals=swap_normals, periodic=periodic, zero_open=zero_open, container=container, fshape=fshape,
                fills=fills)


def bench_config(n, size=None):
    """-> (name, options dict, generator of (corners (nbe,3,3) float32, normals (nbe,3) float32)).  `size` overrides the
    BASELINE size of the family (tests use smaller members)."""
    if n == 1:
        w = config1()
        return w["name"], w, (lambda: (w["tris"], np.ascontiguousarray(w["normals"], dtype=np.float32)))
    if n == 2:
        lv = 3 if size is None else int(size)
        w = spheric2_family(SPHERIC2_DR / 2 ** (lv - 1), 1.25, levels=lv, base_dr=SPHERIC2_DR)
        name = "configs[1]: spheric2 subdivided x%d, dr=%.6g, fshape geometry fill" % (4 ** lv, float(w["dr"]))
        opts = _options(w["dr"], w["fills"], container=w["container"], fshape=w["fshape"])
        return name, opts, (lambda: (np.ascontiguousarray(w["tris"], dtype=np.float32), np.ascontiguousarray(w["normals"], dtype=np.float32)))
    if n == 3:
        nq = 1581 if size is None else int(size)
        dr = 0.01
        L, H = nq * 1.25 * dr, max(6, nq // 4) * 1.25 * dr

        def gen():
            v, t, nn = mg.channel(n=nq, dr=dr)
            return np.ascontiguousarray(v[t], dtype=np.float32), np.ascontiguousarray(nn, dtype=np.float32)
        return ("configs[2]: x/y-periodic channel, %d x %d quads per wall, dr=%g" % (nq, nq, dr),
                _options(dr, [("geometry", (L / 2, L / 2, H / 2), np.float32(dr))], periodic=(True, True, False)), gen)
    if n == 4:
        side = 1000 if size is None else int(size)
        dr = 1.0 / side

        def gen():
            v, t, nn = mg.sphere_in_tank(side_cells=side, dr=dr, hfac=0.5, sphere_levels=7 if side >= 400 else (5 if side >= 100 else 2))
            return np.ascontiguousarray(v[t], dtype=np.float32), np.ascontiguousarray(nn, dtype=np.float32)
        return ("configs[3]: sphere in tank, %d^3-cell lattice, h=dr/2" % (side + 1),
                _options(dr, [("geometry", (4 * dr, 4 * dr, 4 * dr), np.float32(dr))]), gen)
    if n == 5:
        side = 2154 if size is None else int(size)
        dr = 1.0 / side

        def gen():
            v, t, nn = mg.box_tank((0, 0, 0), (1, 1, 1), 0.53 * dr, inward=False)
            return np.ascontiguousarray(v[t], dtype=np.float32), np.ascontiguousarray(nn, dtype=np.float32)
        return ("configs[4]: subdivided tank, %d^3-cell lattice, h=0.53dr" % (side + 1),
                _options(dr, [("geometry", (0.5, 0.5, 0.5), np.float32(dr))]), gen)
    raise ValueError("config must be 1..5")


def bench_config_name(n):
    return {1: "configs[0]: resources/spheric2.ini", 2: "configs[1]: spheric2 subdivided x64 at dr/4, fshape geometry fill",
            3: "configs[2]: x/y-periodic channel, ~1e7 triangles", 4: "configs[3]: sphere in tank, 4.8e7 triangles, 1001^3-cell lattice",
            5: "configs[4]: subdivided tank, 2e8 triangles, 1e10-cell lattice"}[n]


def cpu_sample(n, steps_total=25):
    """The bounded sample of config n that BOTH CPU arms of bench.py time (the cpu_baseline leg of the GPU arm and
    `--impl reference`): the same geometry family, options and mesh refinement relative to dr at a coarser dr, sized so that one
    pass costs seconds on the host cores.  -> (name, options, corners, normals)"""
    small = steps_total > 30
    if n in (1, 2):
        dr = 0.06 if small else 0.045
        w = spheric2_family(dr, 0.625)
        if n == 1:
            w["fills"] = [("box", (0.2, 0.2, 0.2), (0.4, 0.4, 0.4))] + w["fills"]
        name = "spheric2 tank + fshape fill at dr=%g, h=0.625dr" % dr
        opts = _options(w["dr"], w["fills"], container=w["container"], fshape=w["fshape"])
        return name, opts, np.ascontiguousarray(w["tris"], dtype=np.float32), np.ascontiguousarray(w["normals"], dtype=np.float32)
    size = {3: 64 if small else 96, 4: 16 if small else 24, 5: 20 if small else 28}[n]
    name, opts, gen = bench_config(n, size)
    c, nn = gen()
    return name, opts, c, nn
