"""GPU tests of the call a user makes: cx_preprocess_host (whole hot path on host buffers, through the host weld) and the
command line `crixus file.ini` (ini -> STL -> GPU -> .vtu / .h5sph), against the CPU oracle and against the golden
outputs of the reference CUDA binary (tests/golden/)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from crixus_b200 import api, meshgen as mg

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _job_for(c):
    keep = []
    job = api.CxJob()
    tris = np.ascontiguousarray(c["tris"], dtype=np.float32).reshape(-1)
    nrm = np.ascontiguousarray(c["normals"], dtype=np.float32).reshape(-1)
    keep += [tris, nrm]
    job.corners = tris.ctypes.data_as(C.POINTER(C.c_float))
    job.normals = nrm.ctypes.data_as(C.POINTER(C.c_float))
    job.nbe = c["tris"].shape[0]
    job.dr = float(c["dr"])
    job.swap_normals, job.zero_open = int(c["swap_normals"]), int(c["zero_open"])
    for j in range(3):
        job.periodic[j] = int(c["periodic"][j])
    if c["container"] is not None:
        job.use_container = 1
        for j in range(3):
            job.cmin[j], job.cmax[j] = c["container"][0][j], c["container"][1][j]
    if c["fshape"] is not None:
        fc = np.ascontiguousarray(c["fshape"][0], dtype=np.float32).reshape(-1)
        fn = np.ascontiguousarray(c["fshape"][1], dtype=np.float32).reshape(-1)
        keep += [fc, fn]
        job.fs_corners = fc.ctypes.data_as(C.POINTER(C.c_float))
        job.fs_normals = fn.ctypes.data_as(C.POINTER(C.c_float))
        job.fs_nbe = c["fshape"][0].shape[0]
    specs = (api.CxFillSpec * max(len(c["fills"]), 1))()
    for i, f in enumerate(c["fills"]):
        if f[0] == "box":
            specs[i].option = 1
            for j in range(3):
                specs[i].lo[j], specs[i].hi[j] = f[1][j], f[2][j]
        else:
            specs[i].option = 2
            for j in range(3):
                specs[i].seed[j] = f[1][j]
            specs[i].dr_wall = float(f[2]) if f[2] is not None else 0.0
    job.fills, job.nfills = specs, len(c["fills"])
    keep.append(specs)
    sb = [np.ascontiguousarray(x, dtype=np.float32).reshape(-1) for x in c.get("special", ())]
    if sb:
        ptrs = (C.POINTER(C.c_float) * len(sb))(*[x.ctypes.data_as(C.POINTER(C.c_float)) for x in sb])
        cnt = (C.c_int64 * len(sb))(*[x.size // 9 for x in sb])
        job.sb_corners, job.sb_nbe, job.nsb = ptrs, cnt, len(sb)
        keep += sb + [ptrs, cnt]
    return job, keep


def _arr(ptr, shape, dtype):
    n = int(np.prod(shape))
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).reshape(shape).copy() if n else np.zeros(shape, dtype)


@pytest.mark.parametrize("name", ["plate", "cube_jitter", "spheric2_mini", "channel", "sphere_tank", "cube_sb", "plate_sb_stress"])
def test_preprocess_host_vs_oracle(name, oracle, cuda_backend):
    c = cases.case(name)
    o = cases.run(oracle, c)
    ctx = cuda_backend.ctx
    for rep in range(2):                      # second call reuses the context's arenas
        job, keep = _job_for(c)
        ctx.preprocess_host(job)
        nv, nbe = int(job.nvert), int(job.nbe)
        assert nv == o["nvert"] and nbe == o["nbe"]
        np.testing.assert_array_equal(_arr(job.pos, (nv + nbe, 4), np.float32)[:, :3], o["pos"][:, :3])
        np.testing.assert_array_equal(_arr(job.norm, (nbe, 4), np.float32)[:, :3], o["norm"][:, :3])
        np.testing.assert_array_equal(_arr(job.ep, (nbe, 4), np.int32)[:, :3], o["ep"][:, :3])
        np.testing.assert_array_equal(_arr(job.surf, (nbe,), np.float32), o["surf"])
        alive = o["pos"][:nv, 0] > -1e9
        np.testing.assert_array_equal(_arr(job.vol, (nv,), np.float32)[alive], o["vol"][alive])
        np.testing.assert_array_equal(_arr(job.cell_idx, (o["cell_idx"].shape[0],), np.int32), o["cell_idx"])
        np.testing.assert_array_equal(np.array(list(job.dimg)), o["dimg"])
        np.testing.assert_array_equal(_arr(job.fpos, (o["fpos"].shape[0],), np.uint32), o["fpos"])
        assert int(job.nfluid_counted) == o["nfluid"]
        assert int(job.nfluid) == int(np.unpackbits(o["fpos"].view(np.uint8)).sum())
        np.testing.assert_array_equal(np.array(list(job.zstat), dtype=np.float32)[:2], np.asarray(o["zstat"], dtype=np.float32)[:2])
        if c["special"]:
            np.testing.assert_array_equal(_arr(job.sbid, (nv + nbe,), np.int32), o["sbid"])
        else:
            assert not job.sbid
        ctx.L.cx_job_free(C.byref(job))
        assert not job.pos


REC = np.dtype([(n, "<f4") for n in ("x", "y", "z", "nx", "ny", "nz", "vol", "surf")] +
               [(n, "<i4") for n in ("kpar", "kfluid", "kent", "kparmob", "iref", "ep1", "ep2", "ep3")])


@pytest.mark.parametrize("name", ["plate", "spheric2_mini", "channel", "channel_yz", "cube_sb", "plate_sb_stress", "config2_l1"])
def test_device_records_equal_host_records(name, cuda_backend, tmp_path):
    """cx_assemble_records_device + the streamed device writers against the host restatement of crixus.cu:973-1097 /
    crixus.cpp:12-259 (cx_assemble_records, itself checked against numpy in tests/test_frontend.py and, through the CLI, against
    the reference binary's files): same records, byte-identical .vtu and .h5sph files."""
    if name == "config2_l1":      # 1.9e5 triangles, ~1e6 particles: several scan blocks and staging chunks
        from crixus_b200 import workloads as wl
        w = wl.spheric2_family(wl.SPHERIC2_DR / 2, 1.25, levels=1, base_dr=wl.SPHERIC2_DR)
        c = dict(cases.case("spheric2_mini"), tris=np.ascontiguousarray(w["tris"], dtype=np.float32),
                 normals=np.ascontiguousarray(w["normals"], dtype=np.float32), dr=np.float32(w["dr"]),
                 fills=[("geometry", (0.5, 0.5, 0.5), float(w["dr"]))])
    else:
        c = cases.case(name)
    ctx = cuda_backend.ctx
    job, keep = _job_for(c)
    ctx.preprocess_host(job)
    for nfl_idx in (int(job.nfluid), int(job.nfluid) + 1):   # the reference's racy count shifts the vertex indices
        n = ctx.L.cx_assemble_records(C.byref(job), nfl_idx, None, 0)
        host = np.zeros(n, dtype=REC)
        assert ctx.L.cx_assemble_records(C.byref(job), nfl_idx, host.ctypes.data, n) == n
        dev = ctx.assemble_records_device(job, nfl_idx)
        assert dev.shape[0] == n
        assert dev.cpu().numpy().tobytes() == host.tobytes()
    if c["special"]:
        assert host["kent"].max() > 0
    dev = ctx.assemble_records_device(job)
    n = dev.shape[0]
    host = np.frombuffer(dev.cpu().numpy().tobytes(), dtype=REC)
    for fmt, hostfn in (("vtu", ctx.L.cx_write_vtu_records), ("h5sph", ctx.L.cx_write_h5sph_records)):
        a, b = str(tmp_path / ("host." + fmt)), str(tmp_path / ("dev." + fmt))
        assert hostfn(host.ctypes.data, n, a.encode()) == 0
        ctx.write_records_device(dev, b, fmt)
        assert open(a, "rb").read() == open(b, "rb").read(), fmt
    if not c["special"]:
        # the chunked writer for runs whose records do not fit in HBM: tiny chunks here (many of them), same bytes
        for chunk in (7, 1 << 20):
            sfile = str(tmp_path / ("stream%d.h5sph" % chunk))
            assert ctx.stream_h5sph_device(job, sfile, chunk_items=chunk) == n
            assert open(sfile, "rb").read() == open(str(tmp_path / "dev.h5sph"), "rb").read(), "streamed h5sph, chunk %d" % chunk
    ctx.L.cx_job_free(C.byref(job))
    assert not job.d_pos


def test_cli_device_and_host_assembly_write_identical_files(tmp_path):
    """`crixus case.ini` with the default device-side assembly / streamed writers and with CX_HOST_ASSEMBLY=1: identical files,
    unsplit and split, vtu and h5sph."""
    import glob
    name = "cube_sb"
    outs = {}
    for mode in ("device", "host"):
        for fmt in ("vtu", "h5sph"):
            for split in (False, True):
                d = tmp_path / ("%s_%s_%d" % (mode, fmt, split))
                d.mkdir()
                os.environ["CX_HOST_ASSEMBLY"] = "1" if mode == "host" else "0"
                try:
                    _, out = _run_cli(d, name, fmt, split=split)
                finally:
                    os.environ.pop("CX_HOST_ASSEMBLY", None)
                assert ("(%s)" % mode) in out
                files = sorted(glob.glob(str(d / "*.vtu")) + glob.glob(str(d / "*.h5sph")))
                assert len(files) == (5 if split else 1)
                outs[(mode, fmt, split)] = {os.path.basename(f): open(f, "rb").read() for f in files}
    for fmt in ("vtu", "h5sph"):
        for split in (False, True):
            assert outs[("device", fmt, split)] == outs[("host", fmt, split)], (fmt, split)


def _run_cli(tmp_path, name, fmt, split=False):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden import write_ini, write_special
    c = cases.case(name)
    stl = str(tmp_path / (name + ".stl"))
    mg.write_stl(stl, c["tris"], c["normals"])
    fsp = None
    if c["fshape"] is not None:
        fsp = str(tmp_path / (name + "_fs.stl"))
        mg.write_stl(fsp, c["fshape"][0], c["fshape"][1])
    ini = str(tmp_path / (name + ".ini"))
    write_ini(ini, c, stl, fsp, write_special(str(tmp_path), c), split)
    if fmt != "vtu":
        text = open(ini).read().replace("format=vtu", "format=" + fmt)
        open(ini, "w").write(text)
    exe = os.path.join(ROOT, "crixus_b200", "bin", "crixus")
    r = subprocess.run([exe, ini], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return c, r.stdout


def _cmp_particles_with_golden(d, gold, shift):
    from make_golden import canonical_segments
    seg = canonical_segments(d["segment"])
    np.testing.assert_array_equal(d["vertex"]["Points"].astype(np.float32), gold["vert_pos"])
    np.testing.assert_array_equal(d["vertex"]["Volume"], gold["vert_vol"])
    np.testing.assert_array_equal(d["fluid"]["Points"].astype(np.float32), gold["fluid_pos"])
    np.testing.assert_array_equal(seg["Points"].astype(np.float32), gold["seg_pos"])
    np.testing.assert_array_equal(seg["Normal"], gold["seg_norm"])
    np.testing.assert_array_equal(seg["Surface"], gold["seg_surf"])
    np.testing.assert_array_equal(seg["VertexParticle"].astype(np.int64) + shift, gold["seg_ep"])
    if "seg_kent" in gold.files:
        np.testing.assert_array_equal(d["vertex"]["KENT"].astype(np.int32), gold["vert_kent"])
        np.testing.assert_array_equal(seg["KENT"].astype(np.int32), gold["seg_kent"])
        assert (np.diff(d["segment"]["KENT"].astype(np.int64)) >= 0).all()     # grouped by id in the file (crixus.cu:1088-1097)


@pytest.mark.parametrize("name", ["spheric2_mini", "channel", "cube", "cube_sb", "sphere_sb"])
def test_cli_vtu_vs_reference_golden(name, tmp_path):
    """`crixus case.ini` -> case.vtu, compared particle by particle with the file the reference binary wrote."""
    import re
    from vtu import read_vtu, split_particles
    c, out = _run_cli(tmp_path, name, "vtu")
    gold = np.load(os.path.join(HERE, "golden", name + ".npz"))
    d = split_particles(read_vtu(str(tmp_path / (name + ".vtu"))))
    counted = int(re.search(r"Creation of (\d+) fluid particles", str(gold["log"])).group(1))
    shift = counted - gold["fluid_pos"].shape[0]          # the reference's racy extra count: 0 or 1
    assert shift in (0, 1)
    _cmp_particles_with_golden(d, gold, shift)
    assert "Creation of %d fluid particles" % gold["fluid_pos"].shape[0] in out


def test_cli_split_output(tmp_path):
    """[output] split=true with special-boundary grids: the .fluid file and the per-id boundary files together hold exactly
    the records of the unsplit file (crixus.cu:1137-1216)."""
    from vtu import read_vtu
    name = "cube_sb"
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    _run_cli(tmp_path / "a", name, "vtu")
    _, out = _run_cli(tmp_path / "b", name, "vtu", split=True)
    whole = read_vtu(str(tmp_path / "a" / (name + ".vtu")))
    nfl = int((whole["ParticleType"] == 1).sum())
    parts = [read_vtu(str(tmp_path / "b" / (name + ".fluid.vtu")))]
    assert parts[0]["n"] == nfl and " - Fluid particles: %d" % nfl in out
    for k in range(4):
        d = read_vtu(str(tmp_path / "b" / ("%s.boundary.kent%d.vtu" % (name, k))))
        assert (d["KENT"] == k).all() and d["n"] == int(((whole["KENT"] == k) & (whole["ParticleType"] > 1)).sum())
        parts.append(d)
    assert not os.path.exists(str(tmp_path / "b" / (name + ".vtu")))
    for key in ("Volume", "Surface", "KENT", "AbsoluteIndex", "Points", "Normal", "VertexParticle", "ParticleType"):
        cat = np.concatenate([p[key] for p in parts])
        order = np.argsort(np.concatenate([p["AbsoluteIndex"] for p in parts]), kind="stable")
        np.testing.assert_array_equal(cat[order], whole[key])


def test_cli_h5sph_equals_vtu(tmp_path):
    """format=h5sph writes 0.<name>.h5sph (crixus.cpp:268) holding the same records as the .vtu."""
    import h5mini
    from vtu import read_vtu
    name = "spheric2_mini"
    _run_cli(tmp_path, name, "vtu")
    _run_cli(tmp_path, name, "h5sph")
    v = read_vtu(str(tmp_path / (name + ".vtu")))
    h, info = h5mini.read_compound_dataset(str(tmp_path / ("0." + name + ".h5sph")))
    assert info["dims"] == [v["n"]]
    np.testing.assert_array_equal(np.stack([h["Coords_0"], h["Coords_1"], h["Coords_2"]], 1).astype(np.float64), v["Points"])
    np.testing.assert_array_equal(h["Volume"], v["Volume"])
    np.testing.assert_array_equal(h["Surface"], v["Surface"])
    np.testing.assert_array_equal(h["ParticleType"].astype(np.uint32), v["ParticleType"])
    np.testing.assert_array_equal(h["AbsoluteIndex"].astype(np.uint32), v["AbsoluteIndex"])
    np.testing.assert_array_equal(np.stack([h["VertexParticle1"], h["VertexParticle2"], h["VertexParticle3"]], 1).astype(np.uint32),
                                  v["VertexParticle"])
    np.testing.assert_array_equal(np.stack([h["Normal_0"], h["Normal_1"], h["Normal_2"]], 1), v["Normal"])


@pytest.mark.parametrize("name", cases.ALL + ["weld_quirks", "config2_l2"])
def test_device_weld_equals_host_weld(name, cuda_backend):
    """cx_weld_device / cx_bbox_device against cx_weld_host (itself pinned to the reference's remove_duplicate_vertices in
    tests/test_oracle.py): same vertices, same numbering, same ids."""
    import torch
    if name == "weld_quirks":        # near-duplicates inside a cell, across a cell border, negative coordinates
        dr = np.float32(0.1)
        rng = np.random.default_rng(5)
        base = rng.uniform(-0.35, 0.35, size=(400, 3)).astype(np.float32)
        pts = np.concatenate([base, base + np.float32(3e-7), base[:50] + np.float32(2e-6),
                              np.array([[0.19999999, 0.05, 0.05], [0.2, 0.05, 0.05], [-0.05, 0.0, 0.0], [0.05, 0.0, 0.0]], dtype=np.float32)])
        pts = np.concatenate([pts, pts[::-1]])[: (2 * pts.shape[0]) // 3 * 3]
        corners = np.ascontiguousarray(pts, dtype=np.float32)
    elif name == "config2_l2":
        from crixus_b200 import workloads as wl
        w = wl.spheric2_family(wl.SPHERIC2_DR / 2, 1.25, levels=2, base_dr=wl.SPHERIC2_DR)
        corners, dr = np.ascontiguousarray(w["tris"], dtype=np.float32).reshape(-1, 3), np.float32(w["dr"])
    else:
        c = cases.case(name)
        corners, dr = np.ascontiguousarray(c["tris"], dtype=np.float32).reshape(-1, 3), np.float32(c["dr"])
    hv, hid = api.weld_host(corners, dr)
    ctx = cuda_backend.ctx
    pos4, ep4, lo, hi = ctx.weld_device(torch.from_numpy(corners).to(ctx.device), dr)
    np.testing.assert_array_equal(lo, corners.min(0))
    np.testing.assert_array_equal(hi, corners.max(0))
    assert pos4.shape[0] == hv.shape[0]
    np.testing.assert_array_equal(pos4.cpu().numpy()[:, :3], hv)
    np.testing.assert_array_equal(ep4.cpu().numpy()[:, :3], hid)
    assert not ep4[:, 3].any() and not pos4[:, 3].any()


def test_device_weld_overfull_cells(cuda_backend):
    """dr far larger than the mesh spacing puts every corner into one dr-cell (> 20 000 corners): the claim loop of such a cell
    is run by a whole block (k_weld_cells_big) -- no host fallback -- with the host weld's outcome."""
    import torch
    v, t, n = mg.plate(60, 0.01)
    corners = np.ascontiguousarray(v[t], dtype=np.float32).reshape(-1, 3)
    assert corners.shape[0] > 20000
    hv, hid = api.weld_host(corners, np.float32(5.0))
    pos4, ep4, lo, hi = cuda_backend.ctx.weld_device(torch.from_numpy(corners).to(cuda_backend.ctx.device), np.float32(5.0))
    assert pos4.shape[0] == hv.shape[0] == v.shape[0]
    np.testing.assert_array_equal(pos4.cpu().numpy()[:, :3], hv)
    np.testing.assert_array_equal(ep4.cpu().numpy()[:, :3], hid)


def test_cli_equals_reference_binary_on_config1(tmp_path):
    """BASELINE configs[0] at full size, live: resources/spheric2.ini's options on the regenerated SPHERIC-2 mesh (47 736
    triangles) through both command lines -- ours and the reference's own CUDA build (oracle/_ref/Crixus_ref, made by
    oracle/build_ref.sh) -- and the two .vtu files compared particle by particle."""
    ref = os.path.join(ROOT, "oracle", "_ref", "Crixus_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/Crixus_ref not built")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import time_reference_cuda as trc
    from make_golden import write_ini
    c = trc.case_for(0)
    c["name"] = "m"
    stl, fsp, ini = str(tmp_path / "m.stl"), str(tmp_path / "m_fs.stl"), str(tmp_path / "m.ini")
    mg.write_stl(stl, c["tris"], c["normals"])
    mg.write_stl(fsp, c["fshape"][0], c["fshape"][1])
    write_ini(ini, c, stl, fsp)
    res = trc.compare_outputs(str(tmp_path), c, ini)
    assert res["bit_exact"] and res["fluid"] == 104940 and res["segments"] == 47736


def test_triangle_size_check(oracle, cuda_backend):
    """cx_check_triangle_size (the criterion of resources/test-triangle-size.c) against the oracle, exactly."""
    import torch
    v, t, n = mg.spheric2(dr=0.06)
    corners = np.ascontiguousarray(v[t], dtype=np.float32).reshape(-1, 3)
    d = torch.from_numpy(corners).to(cuda_backend.ctx.device)
    for dr in (0.06, 0.05, 0.03):
        got = cuda_backend.ctx.check_triangle_size(d, np.float32(dr))
        ref = oracle.triangle_size(corners, np.float32(dr))
        assert got[0] == ref[0] and np.float32(got[1]) == np.float32(ref[1]) and np.float32(got[2]) == np.float32(ref[2]), (dr, got, ref)
