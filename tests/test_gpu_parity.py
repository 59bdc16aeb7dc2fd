"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on the same inputs.
Integer/index outputs and the fluid bitfield must be bit-exact; float outputs are required to be bit-exact too
(the kernels pin the reference's FMA pattern), the voxel-quantum tolerance is only reported on failure."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _cmp_pipeline(o, g, c):
    nv = o["nvert"]
    assert g["nvert"] == nv and g["nbe"] == o["nbe"]
    np.testing.assert_array_equal(g["verts"], o["verts"])
    np.testing.assert_array_equal(g["epv"], o["epv"])
    assert bytes(g["params_domain"]) == bytes(o["params_domain"])
    # set_bound_elem
    np.testing.assert_array_equal(g["zstat"], o["zstat"])
    np.testing.assert_array_equal(g["pre_ep"][:, :3], o["pre_ep"][:, :3])
    np.testing.assert_array_equal(g["pre_surf"], o["pre_surf"])
    np.testing.assert_array_equal(g["pre_bary"][:, :3], o["pre_bary"][:, :3])
    # binning (stable order on both sides) + periodic relinks
    np.testing.assert_array_equal(g["cell_idx"], o["cell_idx"])
    np.testing.assert_array_equal(g["ep"][:, :3], o["ep"][:, :3])
    np.testing.assert_array_equal(g["pos"][:, :3], o["pos"][:, :3])
    np.testing.assert_array_equal(g["norm"][:, :3], o["norm"][:, :3])
    np.testing.assert_array_equal(g["surf"], o["surf"])
    for d in range(3):
        if c["periodic"][d]:
            np.testing.assert_array_equal(g["newlink%d" % d], o["newlink%d" % d])
    assert g["missing_partners"] == o["missing_partners"] == 0
    np.testing.assert_array_equal(g["trisize"], o["trisize"])
    # volumes: bit-exact
    if "vol" in o:
        alive = o["pos"][:nv, 0] > -1e9
        gv, ov = g["vol"][alive], o["vol"][alive]
        if not np.array_equal(gv, ov):
            q = float(c["dr"] / 40) ** 3
            bad = np.nonzero(gv != ov)[0]
            pytest.fail("%d of %d vertex volumes differ; max |diff| = %.3f voxel quanta; first: %s" %
                        (bad.size, gv.size, np.abs(gv - ov).max() / q, [(int(b), float(gv[b]), float(ov[b])) for b in bad[:5]]))
    if "sbid" in o:
        np.testing.assert_array_equal(g["sbid"], o["sbid"])
    if "fpos" in o:
        assert bytes(g["params_fill"]) == bytes(o["params_fill"])
        np.testing.assert_array_equal(g["dimg"], o["dimg"])
        assert g["fill_counts"] == o["fill_counts"]
        assert g["nfluid"] == o["nfluid"]
        np.testing.assert_array_equal(g["fpos"], o["fpos"])


@pytest.mark.parametrize("name", cases.ALL + cases.EXTRA)
def test_case_vs_oracle(name, oracle, cuda_backend):
    c = cases.case(name)
    o = cases.run(oracle, c)
    g = cases.run(cuda_backend, c)
    _cmp_pipeline(o, g, c)


@pytest.mark.parametrize("name", cases.ALL)
def test_case_vs_reference_golden(name, cuda_backend):
    """CUDA path against outputs of the reference CUDA binary itself (tests/golden/, made by make_golden.py)."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture %s not generated yet" % name)
    gold = np.load(path)
    c = cases.case(name)
    g = cases.run(cuda_backend, c)
    from golden_compare import compare_with_golden
    compare_with_golden(g, gold, c)


@pytest.mark.parametrize("name", ["plate_jitter", "sphere_tank", "channel"])
def test_special_boundary_stress_vs_oracle(name, oracle, cuda_backend):
    """cx_identify_special_boundary (cell-binned candidate lookup, hoisted per-triangle records) against the oracle's
    all-pairs loop on adversarial grids -- every kind, including slivers / zero-area triangles (whole-domain fallback of the
    candidate range) and the rounding-sensitive edge cases: bit-exact."""
    c = cases.case(name)
    for seed in range(4):
        sb = cases.special_stress(c["tris"], seed, n=240)
        c2 = dict(c, special=[sb[:80], sb[80:160], sb[160:]])
        o, g = cases.run(oracle, c2, stages=()), cases.run(cuda_backend, c2, stages=())
        np.testing.assert_array_equal(g["sbid"], o["sbid"])
        assert 0 < (o["sbid"] > 0).sum() < o["sbid"].size


@pytest.mark.parametrize("axis", ["", "0", "1", "2", "chunk", "scan", "overflow"])
@pytest.mark.parametrize("wide", ["0", "1"])
def test_both_vert_volume_kernels(wide, axis):
    """calc_vert_volume has two line kernels (plane_kill<WIDE>, chosen from dr), each in three forms (voxel lines along x, y or
    z, chosen per vertex from its normal; CX_VV_AXIS forces one for all vertices): every one must be bit-exact on every kind of
    mesh: the fine-dr case (several voxels within eps of a plane), the commensurate case and an ordinary one.  "chunk": the
    owned vertices in chunks of 777 (the scratch of the split kernels is re-used from chunk to chunk); "scan": the fan
    gather without the adjacency lists; "overflow": a plane store so small that most fans take the monolithic kernel."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CX_VV_WIDE=wide)
    if axis == "chunk":
        env["CX_VV_CHUNK"] = "777"
    elif axis == "overflow":
        env["CX_VV_PLANE_CAP"] = "100"   # plane store of 100 fan triangles per chunk: most fans overflow into the monolithic kernel
    elif axis == "scan":
        env["CX_VV_SCAN"] = "1"       # fans by scanning the 27 cells instead of the vertex -> segment adjacency
    elif axis:
        env["CX_VV_AXIS"] = axis
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "run_case.py"), "cube_tiny", "cube_h05", "spheric2_mini", "plate_jitter"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
