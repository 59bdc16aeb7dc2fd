"""GPU tests of the error paths: where the reference fails silently on the device (a thread returns, an out-of-bounds write)
or bails out of crixus_main with a return.h code, the C-ABI reports a code and the context stays usable.
  trimax exceeded        crixus_d.cu:306  (thread returns)             -> CX_TRIMAX_EXCEEDED  -102
  fan not in 27 cells    crixus_d.cu:316-369 (reads uninitialised tri) -> CX_FAN_INCOMPLETE   -103
  no periodic partner    crixus.cu:401-409 NO_PER_VERT                 -> CX_NO_PER_VERT      -9
  seed outside lattice   crixus.cu:799-804 (out-of-bounds write)       -> CX_SEED_OUTSIDE     -104
  empty / misplaced box  crixus.cu:759-773 FLUID_NDEF / FLUID_BBOX_TOO_SMALL -> -6 / -17"""
import ctypes as C

import numpy as np
import pytest

import cases
from crixus_b200 import api, meshgen as mg
from test_gpu_pipeline import _arr, _job_for

pytestmark = pytest.mark.gpu


def _fan(k, r=0.07):
    ang = 2 * np.pi * np.arange(k) / k
    v = np.concatenate([[[0.5, 0.5, 0.0]], np.stack([0.5 + r * np.cos(ang), 0.5 + r * np.sin(ang), np.zeros(k)], 1)]).astype(np.float32)
    t = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)], dtype=np.int32)
    c = cases.case("fan20")
    c.update(tris=np.ascontiguousarray(v[t]), normals=np.tile(np.array([0, 0, 1], dtype=np.float32), (k, 1)))
    return c


def _still_works(cuda_backend, oracle):
    c = cases.case("plate")
    g, o = cases.run(cuda_backend, c, stages=("volume",)), cases.run(oracle, c, stages=("volume",))
    np.testing.assert_array_equal(g["vol"], o["vol"])


def test_trimax_exceeded(cuda_backend, oracle):
    g = cases.run(cuda_backend, _fan(50), stages=("volume",))       # exactly trimax = 50 triangles: allowed
    o = cases.run(oracle, _fan(50), stages=("volume",))
    np.testing.assert_array_equal(g["vol"], o["vol"])
    with pytest.raises(api.CxError) as e:
        cases.run(cuda_backend, _fan(51), stages=("volume",))
    assert e.value.code == -102
    _still_works(cuda_backend, oracle)


def test_fan_incomplete_when_triangles_exceed_the_cells(cuda_backend, oracle):
    v, t, n = mg.plate(3, 1.0)                                        # triangles 12 x dr: the mesh precondition is violated
    c = cases.case("plate")
    c.update(tris=np.ascontiguousarray(v[t]), normals=n, container=((0, 0, 0), (3.0, 3.0, 0.2)), fills=[("box", (0.4, 0.4, 0.08), (0.6, 0.6, 0.12))])
    with pytest.raises(api.CxError) as e:
        cases.run(cuda_backend, c, stages=("volume",))
    assert e.value.code == -103
    job, keep = _job_for(c)                                           # the whole-path call reports the precondition, too
    with pytest.raises(api.CxError):
        cuda_backend.ctx.preprocess_host(job)
    assert int(job.tri_failed) > 0 and job.tri_maxd > float(c["dr"])
    _still_works(cuda_backend, oracle)


def test_missing_periodic_partner(cuda_backend, oracle):
    c = cases.case("channel")
    tris = c["tris"].copy()
    L = float(tris[:, :, 0].max())
    on_max = np.isclose(tris[:, :, 0], L) & np.isclose(tris[:, :, 1], tris[:, :, 1].flat[np.argmin(np.abs(tris[:, :, 1] - 0.3))])
    assert on_max.any()
    tris[:, :, 1][on_max] += np.float32(2e-3)                         # one max-face vertex no longer mirrors a min-face vertex
    c["tris"] = tris
    g, o = cases.run(cuda_backend, c, stages=()), cases.run(oracle, c, stages=())
    assert g["missing_partners"] >= 1 and o["missing_partners"] >= 1
    np.testing.assert_array_equal(g["newlink0"], o["newlink0"])       # the vertices that do have a partner are still linked
    job, keep = _job_for(c)
    with pytest.raises(api.CxError) as e:
        cuda_backend.ctx.preprocess_host(job)
    assert e.value.code == -9
    _still_works(cuda_backend, oracle)


def test_fill_input_errors(cuda_backend, oracle):
    ctx = cuda_backend.ctx
    c = cases.case("cube")
    c["fills"] = [("geometry", (5.0, 5.0, 5.0), 0.08)]                # seed outside the lattice
    job, keep = _job_for(c)
    with pytest.raises(api.CxError) as e:
        ctx.preprocess_host(job)
    assert e.value.code == -104
    c = cases.case("plate")
    c["fills"] = [("box", (0.4, 0.4, 0.1), (0.4, 0.6, 0.12))]         # empty box: FLUID_NDEF
    job, keep = _job_for(c)
    with pytest.raises(api.CxError) as e:
        ctx.preprocess_host(job)
    assert e.value.code == -6
    c["fills"] = [("box", (0.4, 0.4, 0.08), (0.6, 0.6, 0.5))]         # box sticks out of the container: FLUID_BBOX_TOO_SMALL
    job, keep = _job_for(c)
    with pytest.raises(api.CxError) as e:
        ctx.preprocess_host(job)
    assert e.value.code == -17
    job = api.CxJob()                                                 # nothing to do
    with pytest.raises(api.CxError) as e:
        ctx.preprocess_host(job)
    assert e.value.code == -11
    _still_works(cuda_backend, oracle)


def test_empty_special_boundary_grid(cuda_backend):
    """a grid without facets tags nothing; ids keep counting (the second grid is #2)"""
    c = cases.case("cube_sb")
    full = cases.run(cuda_backend, c, stages=())["sbid"]
    c2 = dict(c, special=[np.zeros((0, 3, 3), dtype=np.float32)] + list(c["special"]))
    job, keep = _job_for(c2)
    cuda_backend.ctx.preprocess_host(job)
    nv, nbe = int(job.nvert), int(job.nbe)
    got = _arr(job.sbid, (nv + nbe,), np.int32)
    np.testing.assert_array_equal(got, np.where(full > 0, full + 1, 0))
    cuda_backend.ctx.L.cx_job_free(C.byref(job))


def test_unaligned_cell_idx_pointer(cuda_backend, oracle):
    """The C-ABI takes caller-owned device pointers as they are: cell_idx_d that is only 4-byte aligned (the scan inside
    cx_bin_segments vectorises its loads when it can) gives the same prefix sums."""
    import torch
    import cases
    c = cases.case("sphere_tank")
    o = cases.run(oracle, c, stages=())
    ctx, dev = cuda_backend.ctx, cuda_backend.ctx.device
    nv, nbe = o["nvert"], o["nbe"]
    p = o["params_domain"]
    from crixus_b200 import api
    pp = api.CxParams.from_buffer_copy(bytes(p))
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    pre_pos = np.zeros((nv + nbe, 4), dtype=np.float32)
    pre_pos[:nv, :3] = o["verts"]
    pre_pos[nv:] = o["pre_bary"]
    pre_norm = np.zeros((nbe, 4), dtype=np.float32)
    pre_norm[:, :3] = c["normals"]
    ins = [d(pre_pos), d(pre_norm), d(o["pre_ep"]), d(o["pre_surf"])]
    outs = [torch.zeros_like(x) for x in ins]
    buf = torch.zeros(pp.gridn[3] + 2, dtype=torch.int32, device=dev)
    cell_idx = buf[1:]
    assert cell_idx.data_ptr() % 16 == 4
    ctx.bin_segments(pp, ins[0], outs[0], ins[1], outs[1], ins[2], outs[2], ins[3], outs[3], cell_idx, nbe, nv)
    np.testing.assert_array_equal(cell_idx.cpu().numpy(), o["cell_idx"])
