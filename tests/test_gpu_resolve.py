"""GPU: the record-driven resolve of the geometry fill's directed tests (k_resolve_cross + k_resolve_inplane, the default at
fine dr where eps >= dr_wall^2 kills the distance test of /root/reference/src/crixus_d.cu:940) against the cell-driven kernel
(k_resolve_blocks, cx_fill_set_resolve_mode(ctx, 1)) -- the six blocked bit planes must be identical, whole lattice, every scene
of tests/test_resolve_enum.py -- and the fill itself against the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

import cases
from backends import CudaBackend
from crixus_b200 import api
from oracle import orc
from test_resolve_enum import _fuzz_scene, _scene

pytestmark = pytest.mark.gpu

SCENES = ["cube", "cube_jitter", "sphere_tank", "channel", "cube_rot", "cube_tilt", "slivers", "cube_far", "tank_mm"]


def _inputs(B, o):
    d = B._d
    return (d(o["fnorm"]), d(o["fep"]), d(o["fposa"]), int(o["fnbe"]), int(o["cnbe"]), d(o["cell_idx"]))


def _cx(p):
    q = api.CxParams()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
    return q


def _planes(B, o, mode, slab=None, part=None):
    p = _cx(o["params_fill"])
    gn, ge, gp, fnbe, cnbe, ci = _inputs(B, o)
    f0 = torch.zeros(o["fpos"].shape[0], dtype=torch.int32, device=B.dev)
    B.ctx.fill_set_resolve_mode(mode)
    try:
        if part is not None:
            return B.ctx.fill_resolve_part(p, f0, gn, ge, gp, fnbe, cnbe, ci, part[0], part[1]).clone()
        k0, k1 = slab if slab is not None else (0, int(o["dimg"][2]))
        return B.ctx.fill_resolve_slab(p, f0, gn, ge, gp, fnbe, cnbe, ci, k0, k1).clone()
    finally:
        B.ctx.fill_set_resolve_mode(0)


@pytest.mark.parametrize("name", SCENES)
def test_resolve_paths_agree(name):
    c = _scene(name)
    o = cases.run(orc.Oracle(), c, stages=("fill",))
    p = o["params_fill"]
    assert p.eps >= p.dr_wall * p.dr_wall          # the regime of the record-driven kernels
    B = CudaBackend()
    rec = _planes(B, o, 0)
    blk = _planes(B, o, 1)
    assert int(rec.ne(0).sum()) > 0
    assert torch.equal(rec, blk), "%d blocked words differ" % int((rec != blk).sum())
    # the fill through the default path == the oracle's
    g = cases.run(B, c, stages=("fill",))
    assert g["nfluid"] == o["nfluid"] and g["fill_counts"] == o["fill_counts"]
    assert np.array_equal(g["fpos"], o["fpos"])


@pytest.mark.parametrize("seed", range(12))
def test_resolve_paths_agree_fuzz(seed):
    """generated boxes (random size, orientation incl. a few eps off the lattice directions, offset from the origin, mesh width)
    with stray triangles: both paths give the same blocked planes, the fill equals the oracle's"""
    c = _fuzz_scene(seed)
    o = cases.run(orc.Oracle(), c, stages=("fill",))
    B = CudaBackend()
    rec, blk = _planes(B, o, 0), _planes(B, o, 1)
    assert torch.equal(rec, blk), "%d blocked words differ" % int((rec != blk).sum())
    g = cases.run(B, c, stages=("fill",))
    assert g["nfluid"] == o["nfluid"] and np.array_equal(g["fpos"], o["fpos"])


def test_resolve_with_free_surface():
    """fshape triangles (k_resolve_fs) next to the record-driven wall kernels, eps = 0.2 as a relative tolerance"""
    c = cases.case("tank_mm")
    c["fills"] = [("geometry", c["fills"][0][1], 0.4)]
    o = cases.run(orc.Oracle(), c, stages=("fill",))
    p = o["params_fill"]
    assert p.eps >= p.dr_wall * p.dr_wall and o["cnbe"] > 0
    B = CudaBackend()
    assert torch.equal(_planes(B, o, 0), _planes(B, o, 1))
    g = cases.run(B, c, stages=("fill",))
    assert g["nfluid"] == o["nfluid"] and np.array_equal(g["fpos"], o["fpos"])


@pytest.mark.parametrize("name", ["sphere_tank", "channel"])
def test_resolve_shares(name):
    """multi-GPU forms: z-slabs and dealt 3dr-cell chunks are disjoint shares of the same blocked planes"""
    c = _scene(name)
    o = cases.run(orc.Oracle(), c, stages=("fill",))
    B = CudaBackend()
    whole = _planes(B, o, 0)
    dz = int(o["dimg"][2])
    cut = dz // 3
    a, b = _planes(B, o, 0, slab=(0, cut)), _planes(B, o, 0, slab=(cut, dz))
    assert int((a & b).ne(0).sum()) == 0 and torch.equal(a | b, whole)
    parts = [_planes(B, o, 0, part=(3, r)) for r in range(3)]
    acc = torch.zeros_like(whole)
    for q in parts:
        assert int((acc & q).ne(0).sum()) == 0
        acc |= q
    assert torch.equal(acc, whole)
