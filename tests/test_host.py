"""CPU tests of the product's host side: the C-ABI library loads and exports every declared symbol, its host glue
(params, seed, weld) equals the oracle's, mesh generators satisfy the reference's mesh precondition, STL round trip."""
import ctypes
import os
import re

import numpy as np
import pytest

import cases
from crixus_b200 import api, meshgen as mg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(cxlib):
    hdr = open(os.path.join(ROOT, "include", "crixus_b200.h")).read()
    declared = set(re.findall(r"\b(cx_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 20
    raw = ctypes.CDLL(api.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), "libcrixus_b200.so does not export %s" % name
    assert b"sm_100a" in cxlib.cx_version()


def test_struct_layouts_match_header(cxlib):
    assert ctypes.sizeof(api.CxParams) == 4 * (3 + 4 * 3 + 4 + 4 * 2 + 3)
    assert ctypes.sizeof(api.CxFillSpec) == 44


def test_no_gpu_fails_loudly(cxlib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    assert cxlib.cx_create(0, ctypes.byref(h)) == -105   # CX_NO_DEVICE: there is no CPU path
    with pytest.raises(api.CxError):
        api.Context(0)


@pytest.mark.parametrize("name", ["plate", "spheric2_mini", "channel", "sphere_tank"])
def test_host_glue_equals_oracle(name, oracle, cxlib):
    c = cases.case(name)
    corners = c["tris"].reshape(-1, 3)
    lo, hi = corners.min(0), corners.max(0)
    p, q = api.params_domain(c["dr"], lo, hi), oracle.params_domain(c["dr"], lo, hi)
    assert bytes(p) == bytes(q)
    for j in range(3):
        p.per[j] = q.per[j] = int(c["periodic"][j])
    api.params_fill_eps(p)
    oracle.params_fill_eps(q)
    if c["container"] is not None:
        api.params_container(p, True, *c["container"])
        oracle.params_container(q, True, *c["container"])
    else:
        api.params_container(p, False)
        oracle.params_container(q, False)
    assert bytes(p) == bytes(q)
    np.testing.assert_array_equal(api.fill_dims(p), oracle.fill_dims(q))
    for s in [(0.5, 0.5, 0.5), (0.1, 0.2, 0.05), (-1.0, 0.0, 0.0)]:
        assert api.seed_index(p, s) == oracle.seed_index(q, s)
    a, b = api.weld_host(corners, c["dr"]), oracle.weld(corners, c["dr"])
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


def test_meshes_satisfy_triangle_size_precondition():
    for name in cases.ALL:
        if name in ("plate_jitter", "plate_sb_stress"):   # deliberately generic: the SURVEY D.4 sensitivity mesh (h=1.25dr plus 10 % jitter)
            continue
        c = cases.case(name)
        t = c["tris"].astype(np.float64)
        bary = t.mean(1, keepdims=True)
        assert np.sqrt(((t - bary) ** 2).sum(-1)).max() <= float(c["dr"]) * (1 - 1e-4), name


def test_stl_roundtrip_and_ascii_rejection(tmp_path):
    v, t, n = mg.spheric2_fshape()
    p = tmp_path / "fs.stl"
    mg.write_stl(str(p), v[t], n)
    assert os.path.getsize(p) == 84 + 50 * 4 == 284          # same size as the shipped resources/spheric2_fshape.stl
    tt, nn = mg.read_stl(str(p))
    np.testing.assert_array_equal(tt, v[t])
    np.testing.assert_array_equal(nn, n)
    q = tmp_path / "ascii.stl"
    q.write_bytes(b"solid x\nendsolid x\n".ljust(100))
    with pytest.raises(ValueError):
        mg.read_stl(str(q))


def test_subdivide_is_watertight_and_scales():
    v, t, n = mg.box_tank((0, 0, 0), (1, 1, 1), 0.25)
    v2, t2, n2 = mg.subdivide(v, t, n, 2)
    assert t2.shape[0] == 16 * t.shape[0]
    e = np.sort(np.concatenate([t2[:, [0, 1]], t2[:, [1, 2]], t2[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).all()


def test_ctypes_binding_covers_the_header(cxlib):
    """crixus_b200/api.py declares argument types for every entry point of include/crixus_b200.h, and the ctypes mirror of
    cx_job has the library's size (api.lib() checks it on load)."""
    hdr = open(os.path.join(ROOT, "include", "crixus_b200.h")).read()
    declared = set(re.findall(r"\b(cx_[a-z_0-9]+)\s*\(", hdr))
    missing = sorted(declared - set(cxlib._declared))
    assert not missing, "no ctypes signature for %s" % missing
    assert cxlib.cx_job_sizeof() == ctypes.sizeof(api.CxJob)


def test_null_context_is_rejected_without_a_gpu(cxlib):
    """The stage entry points added for the rows either side of the hot path validate their arguments before touching the
    device: CX_WRONG_INPUT (-11), never a crash."""
    z = ctypes.c_void_p()
    n = ctypes.c_int64(0)
    p = api.CxParams()
    d = (ctypes.c_int64 * 3)(1, 1, 1)
    assert cxlib.cx_identify_special_boundary(z, ctypes.byref(p), z, z, z, 0, 0, z, z, 0, z, 1) == -11
    assert cxlib.cx_assemble_records_device(z, ctypes.byref(p), d, z, z, z, z, z, z, z, 0, 0, 0, z, 0, ctypes.byref(n)) == -11
    assert cxlib.cx_write_vtu_records_device(z, z, 0, b"/tmp/x.vtu") == -11
    assert cxlib.cx_write_h5sph_records_device(z, z, 0, b"/tmp/x.h5sph") == -11
    assert cxlib.cx_write_split_records(z, 3, 0, 1, b"x", 1) == -11
    assert cxlib.cx_write_split_records(z, 0, 0, 0, b"x", 1) == -11
    assert cxlib.cx_fill_geometry_slab(z, ctypes.byref(p), z, z, z, z, 0, 0, z, 0, 0, 1, 2, ctypes.byref(n), None) == -11
    assert cxlib.cx_fill_set_resolve_mode(z, 0) == -11


def test_host_glue_fuzz_against_oracle(oracle, cxlib):
    """cx_params_domain / _fill_eps / _container / cx_fill_dims / cx_seed_index (csrc/ctx.cu) against the oracle's independent
    restatement of crixus.cu:262-283,464-467,698-722,799-804 on 2000 random domains, spacings, containers and seeds:
    identical bytes / integers (float-to-int truncation, (extent+eps)/gridn, round() of a double expression)."""
    rng = np.random.default_rng(42)
    for it in range(2000):
        scale = 10 ** rng.uniform(-3, 2)
        lo = (rng.uniform(-1, 1, 3) * scale).astype(np.float32)
        ext = (10 ** rng.uniform(-1.5, 0.5, 3) * scale).astype(np.float32)
        hi = (lo + ext).astype(np.float32)
        dr = np.float32(float(ext.min()) * 10 ** rng.uniform(-1.8, 0.3))
        p, q = api.params_domain(dr, lo, hi), oracle.params_domain(dr, lo, hi)
        assert bytes(p) == bytes(q), (it, lo, hi, dr)
        for j in range(3):
            p.per[j] = q.per[j] = int(rng.integers(0, 2))
        api.params_fill_eps(p)
        oracle.params_fill_eps(q)
        if it % 2:
            cmin = (lo + rng.uniform(-0.2, 0.5, 3).astype(np.float32) * ext).astype(np.float32)
            cmax = (cmin + rng.uniform(0.05, 1.2, 3).astype(np.float32) * ext).astype(np.float32)
            api.params_container(p, True, cmin, cmax)
            oracle.params_container(q, True, cmin, cmax)
        else:
            api.params_container(p, False)
            oracle.params_container(q, False)
        assert bytes(p) == bytes(q), (it, "container")
        np.testing.assert_array_equal(api.fill_dims(p), oracle.fill_dims(q))
        for _ in range(3):
            s = (lo + rng.uniform(-0.3, 1.3, 3).astype(np.float32) * ext).astype(np.float32)
            assert api.seed_index(p, s) == oracle.seed_index(q, s), (it, s)
