"""CPU check of the culling rules of the record-driven fill resolve (crixus_b200/csrc/fill_enum.h -- the header fill.cu compiles
for the device) against the oracle's restatement of checkCollision (/root/reference/src/crixus_d.cu:753-857): on seeded scenes
with a dead distance test (eps >= dr_wall^2, the regime the record-driven kernels run in) every (lattice cell, direction,
segment) triple the reference's arithmetic blocks must be among the enumerated candidates.  tests/resolve_enum_check.c does the
record-by-record comparison; the GPU tests compare the blocked planes the kernels produce with the cell-driven kernel and the
oracle (tests/test_gpu_parity.py::test_resolve_paths_agree)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cases
from crixus_b200 import meshgen as mg
from oracle import orc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libresolve_enum.so")
    srcs = [os.path.join(HERE, "resolve_enum_check.c"), os.path.join(HERE, "..", "crixus_b200", "csrc", "fill_enum.h"),
            os.path.join(HERE, "..", "oracle", "crixus_oracle.c")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["gcc", "-std=c11", "-O2", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-Wall",
                               "-Wno-unused-function", srcs[0], "-o", so, "-lm"])
    L = C.CDLL(so)
    L.rec_enum_check.restype = C.c_int64
    L.rec_enum_check.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    L.plane_cache_check.restype = C.c_int64
    L.plane_cache_check.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_void_p]
    return L


def _rot(ax, ang):
    c, s = np.cos(ang), np.sin(ang)
    m = np.eye(3)
    a, b = [(1, 2), (2, 0), (0, 1)][ax]
    m[a, a] = c; m[a, b] = -s; m[b, a] = s; m[b, b] = c
    return m


def _scene(name):
    """-> case dict with a geometry fill whose dr_wall makes the distance test dead"""
    if name == "cube":               # walls in lattice planes: the "ray in plane" rule on every wall cell
        c = cases.case("cube_h05")
    elif name == "cube_jitter":      # generic planes
        c = cases.case("cube_jitter")
    elif name == "sphere_tank":      # curved, oblique facets
        c = cases.case("sphere_tank")
    elif name == "channel":          # periodic grid (wrapped neighbour cells)
        c = cases.case("channel")
    elif name == "cube_rot":         # a box rotated out of every lattice direction, its walls a hair off the axes for one of them
        v, t, n = mg.box_tank((0, 0, 0), (1, 1, 1), 0.08, inward=True)
        R = _rot(2, 0.31) @ _rot(0, 2e-6) @ _rot(1, 0.17)
        v = (v.astype(np.float64) @ R.T).astype(np.float32)
        n = mg.geometric_normals(v, t)
        n[(n * (v[t].mean(1) - v.mean(0))).sum(1) > 0] *= -1
        c = dict(cases.case("cube"), tris=np.ascontiguousarray(v[t]), normals=np.ascontiguousarray(n), dr=np.float32(0.07))
        c["fills"] = [("geometry", tuple(float(x) for x in v.mean(0)), 0.07)]
    elif name == "cube_tilt":        # walls within eps of the lattice planes but not in them (|b| around eps)
        v, t, n = mg.box_tank((0, 0, 0), (1, 1, 1), 0.1, inward=True)
        R = _rot(0, 1.5e-4) @ _rot(1, -0.9e-4)
        v = (v.astype(np.float64) @ R.T).astype(np.float32)
        n = mg.geometric_normals(v, t)
        n[(n * (v[t].mean(1) - v.mean(0))).sum(1) > 0] *= -1
        c = dict(cases.case("cube"), tris=np.ascontiguousarray(v[t]), normals=np.ascontiguousarray(n), dr=np.float32(0.08))
    elif name == "slivers":          # triangles near the well-conditioned threshold, large against dr, random orientations
        rng = np.random.default_rng(11)
        tr = []
        for _ in range(400):
            o = rng.uniform(0.2, 0.8, 3)
            a = rng.normal(size=3); a /= np.linalg.norm(a)
            b = np.cross(a, rng.normal(size=3)); b /= np.linalg.norm(b)
            la, lb = rng.uniform(0.02, 0.25, 2)
            ang = rng.choice([0.11, 0.2, 0.7, 1.4, 2.9])    # sin^2 = 0.012 ... 1
            tr.append([o, o + la * a, o + lb * (np.cos(ang) * a + np.sin(ang) * b)])
        tr = np.array(tr, dtype=np.float32)
        box_v, box_t, box_n = mg.box_tank((0, 0, 0), (1, 1, 1), 0.1, inward=True)
        n = mg.geometric_normals(tr.reshape(-1, 3), np.arange(3 * len(tr), dtype=np.int32).reshape(-1, 3))
        c = dict(cases.case("cube"), tris=np.ascontiguousarray(np.concatenate([box_v[box_t], tr])),
                 normals=np.ascontiguousarray(np.concatenate([box_n, n])), dr=np.float32(0.05))
    elif name == "cube_far":         # far from the coordinate origin: one ulp of a coordinate is 25 x eps
        c = cases.case("cube_h05")   # (walls "in" lattice planes that the rounded lattice positions straddle)
        c["tris"] = np.ascontiguousarray(c["tris"] + np.array([3000.0, -2000.0, 1000.0], dtype=np.float32))
        c["fills"] = [("geometry", (3000.5, -1999.5, 1000.5), 0.1)]
    elif name == "tank_mm":          # eps = 0.2 as a relative tolerance of the segment test
        c = cases.case("tank_mm")
        c["fshape"] = None
    else:
        raise KeyError(name)
    c = dict(c)
    ext = float((c["tris"].reshape(-1, 3).max(0) - c["tris"].reshape(-1, 3).min(0)).max())
    drw = 0.9 * np.sqrt(1e-5 * ext)
    fills = []
    for f in c["fills"]:
        fills.append(("geometry", f[1], float(drw)) if f[0] == "geometry" else f)
    c["fills"] = fills
    return c


def _fuzz_scene(seed):
    """a box of random size, orientation (every third one within a few eps of the lattice directions), position and mesh width,
    plus a few free-floating triangles; dr and dr_wall drawn so that the distance test is dead"""
    rng = np.random.default_rng(9000 + seed)
    dr = float(rng.choice([0.05, 0.07, 0.1]))
    h = dr * float(rng.choice([0.5, 0.8, 1.25]))
    size = np.array([rng.integers(6, 12) * h for _ in range(3)])
    v, t, n = mg.box_tank((0, 0, 0), tuple(size), h, inward=True)
    if seed % 3 == 0:
        ang = rng.uniform(-3e-4, 3e-4, 3)
    elif seed % 3 == 1:
        ang = rng.uniform(-0.6, 0.6, 3)
    else:
        ang = np.array([0.0, 0.0, rng.uniform(-0.5, 0.5)])     # z stays a lattice direction: capable walls off the lattice planes
    R = _rot(0, ang[0]) @ _rot(1, ang[1]) @ _rot(2, ang[2])
    shift = rng.uniform(-1, 1, 3) * float(rng.choice([0.0, 3.0, 400.0]))
    v = (v.astype(np.float64) @ R.T + shift).astype(np.float32)
    nrm = mg.geometric_normals(v, t)
    nrm[(nrm * (v[t].mean(1) - v.mean(0))).sum(1) > 0] *= -1
    tr = [v[t]]
    nn = [nrm]
    k = int(rng.integers(0, 12))
    if k:
        o = v.mean(0) + rng.uniform(-0.2, 0.2, (k, 3)) * size
        a = rng.normal(size=(k, 3)); b = rng.normal(size=(k, 3))
        extra = np.stack([o, o + a * h * rng.uniform(0.3, 1.5, (k, 1)), o + b * h * rng.uniform(0.3, 1.5, (k, 1))], 1).astype(np.float32)
        tr.append(extra)
        nn.append(mg.geometric_normals(extra.reshape(-1, 3), np.arange(3 * k, dtype=np.int32).reshape(-1, 3)))
    tris = np.ascontiguousarray(np.concatenate(tr), dtype=np.float32)
    ext = float((tris.reshape(-1, 3).max(0) - tris.reshape(-1, 3).min(0)).max())
    c = dict(cases.case("cube"), tris=tris, normals=np.ascontiguousarray(np.concatenate(nn), dtype=np.float32), dr=np.float32(dr))
    c["fills"] = [("geometry", tuple(float(x) for x in v.mean(0)), float(0.9 * np.sqrt(1e-5 * ext)))]
    return c


@pytest.mark.parametrize("seed", range(12))
def test_candidates_cover_reference_fuzz(lib, seed):
    c = _fuzz_scene(seed)
    o = cases.run(orc.Oracle(), c, stages=("fill",))
    p = o["params_fill"]
    assert p.eps >= p.dr_wall * p.dr_wall
    nwall = int(o["fnbe"] - o["cnbe"])
    out = np.zeros(8, dtype=np.int64)
    pos, norm, ep, ci = (np.ascontiguousarray(o[k]) for k in ("fposa", "fnorm", "fep", "cell_idx"))
    assert lib.rec_enum_check(C.byref(p), pos.ctypes.data, norm.ctypes.data, ep.ctypes.data, nwall, ci.ctypes.data, out.ctypes.data, 1) == 0
    print("fuzz %d: %d records (%d eligible), reference blocks %d triples; candidates: %d crossing, %d in-plane" % (seed, out[5], out[4], out[0], out[2], out[3]))
    assert out[5] == nwall and out[4] > 0
    assert out[1] == 0


@pytest.mark.parametrize("name", ["cube", "cube_jitter", "sphere_tank", "channel", "cube_rot", "cube_tilt", "slivers", "cube_far", "tank_mm"])
def test_candidates_cover_reference(lib, name):
    c = _scene(name)
    B = orc.Oracle()
    o = cases.run(B, c, stages=("fill",))
    p = o["params_fill"]
    assert p.eps >= p.dr_wall * p.dr_wall
    nwall = int(o["fnbe"] - o["cnbe"])
    out = np.zeros(8, dtype=np.int64)
    pos, norm, ep, ci = (np.ascontiguousarray(o[k]) for k in ("fposa", "fnorm", "fep", "cell_idx"))
    rc = lib.rec_enum_check(C.byref(p), pos.ctypes.data, norm.ctypes.data, ep.ctypes.data, nwall, ci.ctypes.data, out.ctypes.data, 1)
    assert rc == 0
    blocked, missed, ncross, ninpl, nelig, nrec, nunsafe, looked = (int(x) for x in out)
    print("%s: %d records (%d eligible, %d not safe), reference blocks %d triples of %d looked at; candidates: %d crossing, %d in-plane"
          % (name, nrec, nelig, nunsafe, blocked, looked, ncross, ninpl))
    assert nrec == nwall
    assert blocked > 0 and nelig > 0
    assert missed == 0


def _layered_scene(offset, far):
    """a z-wall in a lattice plane and copies of it moved along the normal by fractions of eps (eps = 1e-5 x extent): planes with
    the same normal bits whose offsets differ by less than the tolerance of the remembered plane -- and by more"""
    v, t, n = mg.box_tank((0, 0, 0), (1, 1, 1), 0.1, inward=True)
    tris, nrm = v[t], n
    floor = np.abs(tris[:, :, 2]).max(1) < 1e-6
    eps = 1e-5 * 1.0
    layers = [tris]
    norms = [nrm]
    for f in offset:
        lay = tris[floor].copy()
        lay[:, :, 2] += np.float32(f * eps)
        layers.append(lay)
        norms.append(nrm[floor])
    tris = np.concatenate(layers).astype(np.float32)
    if far:
        tris = tris + np.array([300.0, -200.0, 100.0], dtype=np.float32)
    c = dict(cases.case("cube"), tris=np.ascontiguousarray(tris), normals=np.ascontiguousarray(np.concatenate(norms), dtype=np.float32),
             dr=np.float32(0.1))
    ctr = tris.reshape(-1, 3).mean(0)
    c["fills"] = [("geometry", tuple(float(x) for x in ctr), float(0.9 * np.sqrt(1e-5)))]
    return c


@pytest.mark.parametrize("far", [False, True])
def test_remembered_plane_is_sound(lib, far):
    """k_resolve_inplane skips a segment whose plane is within eps / 2 (3 eps / 4 through the per-cell summaries) of a finished plane
    with the same normal bits: every cell the reference puts in the skipped segment's plane must be in the finished plane's
    2-eps mask -- also far from the coordinate origin, where one ulp of a coordinate is a third of eps"""
    c = _layered_scene([0.2, 0.45, 0.7, 1.1, 2.3, -0.3, -0.74], far)
    o = cases.run(orc.Oracle(), c, stages=("fill",))
    p = o["params_fill"]
    nwall = int(o["fnbe"] - o["cnbe"])
    pos, norm, ep, ci = (np.ascontiguousarray(o[k]) for k in ("fposa", "fnorm", "fep", "cell_idx"))
    for tol in (0.5, 0.75):
        out = np.zeros(3, dtype=np.int64)
        assert lib.plane_cache_check(C.byref(p), pos.ctypes.data, norm.ctypes.data, ep.ctypes.data, nwall, ci.ctypes.data, tol, out.ctypes.data) == 0
        print("far=%s tol=%.2f: %d plane pairs, %d in-plane cells, %d outside the other mask" % (far, tol, out[0], out[1], out[2]))
        assert out[0] > 0 and out[1] > 0
        assert out[2] == 0
    # the check has teeth: with a tolerance of 3 eps the layer at 2.3 eps would vouch for the one at 0.2 eps, and the inclusion fails
    out = np.zeros(3, dtype=np.int64)
    lib.plane_cache_check(C.byref(p), pos.ctypes.data, norm.ctypes.data, ep.ctypes.data, nwall, ci.ctypes.data, 3.0, out.ctypes.data)
    assert out[2] > 0
