"""The reference's process surface, CPU only: .ini reader against the reference's own INIReader (vendored inih, linked
into oracle/_ref/libcrixus_refshim.so), VTU writer byte-for-byte against the reference's own vtk_output
(src/crixus.cpp:62-259), the hand-written HDF5 writer against an independent parser (tests/h5mini.py; there is no
libhdf5 in this image), binary-STL reader edge cases (src/crixus.cu:126-221)."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from crixus_b200 import api, meshgen as mg

REC = np.dtype([(n, "<f4") for n in ("x", "y", "z", "nx", "ny", "nz", "vol", "surf")] +
               [(n, "<i4") for n in ("kpar", "kfluid", "kent", "kparmob", "iref", "ep1", "ep2", "ep3")])
H5_NAMES = ["Coords_0", "Coords_1", "Coords_2", "Normal_0", "Normal_1", "Normal_2", "Volume", "Surface", "ParticleType",
            "FluidType", "KENT", "MovingBoundary", "AbsoluteIndex", "VertexParticle1", "VertexParticle2", "VertexParticle3"]


def _records(n, seed=1):
    rng = np.random.default_rng(seed)
    r = np.zeros(n, dtype=REC)
    for k in ("x", "y", "z", "nx", "ny", "nz", "vol", "surf"):
        r[k] = rng.standard_normal(n).astype(np.float32)
    r["kpar"] = rng.integers(1, 4, n)
    r["kfluid"] = 1
    r["iref"] = np.arange(n)
    for k in ("ep1", "ep2", "ep3"):
        r[k] = rng.integers(0, max(n, 1), n)
    return r


INI_TRICKY = """; comment line
# another comment
[mesh]
stlfile = tank.stl   ; inline comment after whitespace
dr=0.018333
swap_normals : yes
zeroOpen=OFF
name with spaces = a;b ;c
long = first
  second line
  third ; x
[Periodicity]
X=TRUE
y = 0
z = maybe
[fill_0]
option=geometry
xseed = 1e-1
yseed = .5abc
zseed = abc
[output]
format=h5sph
name = run 7
[system]
gpu-id = 0x1
bad line without separator
[unterminated
k=v
"""


def _write(tmp_path, text, name="c.ini"):
    p = tmp_path / name
    p.write_bytes(text.encode())
    return str(p)


def test_ini_reader_equals_reference(tmp_path, cxlib, refshim):
    R = refshim.R
    R.refshim_ini_open.restype = C.c_void_p
    R.refshim_ini_open.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    R.refshim_ini_close.argtypes = [C.c_void_p]
    R.refshim_ini_get.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    R.refshim_ini_get_integer.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_long]
    R.refshim_ini_get_integer.restype = C.c_long
    R.refshim_ini_get_real.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_double]
    R.refshim_ini_get_real.restype = C.c_double
    R.refshim_ini_get_boolean.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    long_line = "[s]\nk=" + "x" * 300 + "\nafter=1\n"
    bom = "﻿[mesh]\ndr=1\n"
    shipped = open("/root/reference/resources/spheric2.ini").read() if os.path.exists("/root/reference/resources/spheric2.ini") else ""
    for idx, text in enumerate([INI_TRICKY, long_line, bom, "", "[a]\nb=1\nb=2\n", shipped]):
        path = _write(tmp_path, text, "c%d.ini" % idx).encode()
        err = C.c_int(0)
        r = R.refshim_ini_open(path, C.byref(err))
        h = api.VP()
        mine = cxlib.cx_ini_open(path, C.byref(h))
        assert mine == err.value, (idx, mine, err.value)
        keys = [("mesh", "stlfile"), ("MESH", "DR"), ("mesh", "swap_normals"), ("mesh", "zeroopen"), ("mesh", "name with spaces"),
                ("mesh", "long"), ("periodicity", "x"), ("periodicity", "y"), ("periodicity", "z"), ("fill_0", "option"),
                ("fill_0", "xseed"), ("fill_0", "yseed"), ("fill_0", "zseed"), ("output", "format"), ("output", "name"),
                ("system", "gpu-id"), ("unterminated", "k"), ("system", "k"), ("s", "k"), ("s", "after"), ("a", "b"), ("nope", "nope"),
                ("fluid_container", "use"), ("fluid_container", "xmax"), ("fill_1", "dr_wall"), ("fill_1", "option")]
        for sec, key in keys:
            a, b = C.create_string_buffer(1024), C.create_string_buffer(1024)
            la = R.refshim_ini_get(r, sec.encode(), key.encode(), b"DEF", a, 1024)
            lb = cxlib.cx_ini_get(h, sec.encode(), key.encode(), b"DEF", b, 1024)
            assert (la, a.value) == (lb, b.value), (idx, sec, key, a.value, b.value)
            assert R.refshim_ini_get_integer(r, sec.encode(), key.encode(), -7) == cxlib.cx_ini_get_integer(h, sec.encode(), key.encode(), -7)
            ra, rb = R.refshim_ini_get_real(r, sec.encode(), key.encode(), -7.5), cxlib.cx_ini_get_real(h, sec.encode(), key.encode(), -7.5)
            assert ra == rb, (sec, key, ra, rb)
            for d in (0, 1):
                assert R.refshim_ini_get_boolean(r, sec.encode(), key.encode(), d) == cxlib.cx_ini_get_boolean(h, sec.encode(), key.encode(), d)
        R.refshim_ini_close(r)
        cxlib.cx_ini_close(h)
    # a file that does not exist: ParseError -1 on both sides
    err = C.c_int(0)
    r = R.refshim_ini_open(b"/nonexistent.ini", C.byref(err))
    h = api.VP()
    assert cxlib.cx_ini_open(b"/nonexistent.ini", C.byref(h)) == err.value == -1
    R.refshim_ini_close(r)
    cxlib.cx_ini_close(h)


@pytest.mark.parametrize("n", [0, 1, 5, 1000])
def test_vtu_writer_is_byte_identical_to_reference(n, tmp_path, cxlib, refshim):
    rec = _records(n)
    a, b = str(tmp_path / "a.vtu"), str(tmp_path / "b.vtu")
    refshim.R.refshim_vtk_output.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    assert refshim.R.refshim_vtk_output(rec.ctypes.data, n, a.encode()) == 0
    assert cxlib.cx_write_vtu_records(rec.ctypes.data, n, b.encode()) == 0
    assert open(a, "rb").read() == open(b, "rb").read()
    if n:
        from vtu import read_vtu
        d = read_vtu(b)
        assert d["n"] == n
        np.testing.assert_array_equal(d["Volume"], rec["vol"])
        np.testing.assert_array_equal(d["Points"], np.stack([rec["x"], rec["y"], rec["z"]], 1).astype(np.float64))
        np.testing.assert_array_equal(d["VertexParticle"], np.stack([rec["ep1"], rec["ep2"], rec["ep3"]], 1).astype(np.uint32))


@pytest.mark.parametrize("n", [0, 1, 7, 4096])
def test_h5sph_writer_against_independent_parser(n, tmp_path, cxlib):
    import h5mini
    rec = _records(n, seed=3)
    path = str(tmp_path / "0.case.h5sph")
    assert cxlib.cx_write_h5sph_records(rec.ctypes.data, n, path.encode()) == 0
    arr, info = h5mini.read_compound_dataset(path, "Compound")       # dataset name: src/crixus.h:21
    assert info["dims"] == [n] and info["record"] == 64
    # member names, order, offsets and types of src/crixus.cpp:21-38
    assert [m[0] for m in info["members"]] == H5_NAMES
    assert [m[1] for m in info["members"]] == [4 * i for i in range(16)]
    assert [m[2] for m in info["members"]] == [np.dtype("<f4")] * 8 + [np.dtype("<i4")] * 8
    assert arr.tobytes() == rec.tobytes()
    assert info["data_address"] % 8 == 0


def test_h5sph_bytes_field_by_field_against_the_format_specification(tmp_path, cxlib):
    """A second, parser-free check of the hand-written HDF5 writer: every metadata field of a 2-record file at its byte offset,
    with the value the HDF5 File Format Specification (version 2.0, superblock version 0) prescribes -- section numbers below.
    (No libhdf5 / h5py in this image; tests/h5mini.py is the structural reader, this is the hex-level one.)"""
    import struct
    rec = _records(2, seed=5)
    path = str(tmp_path / "0.two.h5sph")
    assert cxlib.cx_write_h5sph_records(rec.ctypes.data, 2, path.encode()) == 0
    b = open(path, "rb").read()
    u8 = lambda o: b[o]                                     # noqa: E731
    u16 = lambda o: struct.unpack_from("<H", b, o)[0]       # noqa: E731
    u32 = lambda o: struct.unpack_from("<I", b, o)[0]       # noqa: E731
    u64 = lambda o: struct.unpack_from("<Q", b, o)[0]       # noqa: E731
    UNDEF = 0xFFFFFFFFFFFFFFFF
    # II.A superblock, version 0
    assert b[0:8] == bytes([0x89]) + b"HDF\r\n\x1a\n"                          # format signature
    assert [u8(8), u8(9), u8(10), u8(11), u8(12)] == [0, 0, 0, 0, 0]           # versions: superblock, free space, root group, reserved, shared header
    assert (u8(13), u8(14), u8(15)) == (8, 8, 0)                                # size of offsets, size of lengths, reserved
    assert (u16(16), u16(18)) == (4, 16)                                        # group leaf node K, group internal node K
    assert u32(20) == 0                                                         # file consistency flags
    assert u64(24) == 0 and u64(32) == UNDEF and u64(48) == UNDEF               # base address, free-space info, driver info
    assert u64(40) == len(b)                                                    # end-of-file address
    # root group symbol table entry (III.C): link name offset, object header address, cache type 1, scratch = B-tree + heap addresses
    root_hdr, btree, heap = u64(64), u64(80), u64(88)
    assert u64(56) == 0 and u32(72) == 1 and u32(76) == 0
    # IV.A.1.a root object header, version 1: one message of type 0x0011 (symbol table) carrying the same two addresses
    assert (u8(root_hdr), u16(root_hdr + 2), u32(root_hdr + 4)) == (1, 1, 1)
    assert u16(root_hdr + 16) == 0x0011 and u16(root_hdr + 18) == 16
    assert (u64(root_hdr + 24), u64(root_hdr + 32)) == (btree, heap)
    # III.D local heap: signature, version, data segment size, free-list head, data address; the name lives at offset 8
    assert b[heap:heap + 4] == b"HEAP" and u8(heap + 4) == 0
    heap_data = u64(heap + 24)
    assert b[heap_data + 8:heap_data + 17] == b"Compound\0"                     # dataset name (src/crixus.h:21)
    free = u64(heap + 16)
    assert u64(heap_data + free) == 1 and u64(heap_data + free + 8) == u64(heap + 8) - free   # one free block: next = 1 (none), size = the rest
    # III.A.1 version-1 B-tree node: group node (type 0), leaf (level 0), one entry, no siblings, key 1 = heap offset of the largest name
    assert b[btree:btree + 4] == b"TREE" and (u8(btree + 4), u8(btree + 5), u16(btree + 6)) == (0, 0, 1)
    assert u64(btree + 8) == UNDEF and u64(btree + 16) == UNDEF
    snod = u64(btree + 32)
    assert u64(btree + 24) == 0 and u64(btree + 40) == 8
    # III.B symbol table node: version 1, one symbol; its entry: name offset 8, object header address, cache type 0
    assert b[snod:snod + 4] == b"SNOD" and (u8(snod + 4), u16(snod + 6)) == (1, 1)
    dset = u64(snod + 16)
    assert u64(snod + 8) == 8 and u32(snod + 24) == 0
    # dataset object header, version 1, four messages: dataspace 0x0001, datatype 0x0003, fill value 0x0005, layout 0x0008
    assert (u8(dset), u16(dset + 2), u32(dset + 4)) == (1, 4, 1)
    o, seen = dset + 16, {}
    for _ in range(4):
        t, sz = u16(o), u16(o + 2)
        seen[t] = o + 8
        o += 8 + sz
        assert sz % 8 == 0                                                      # version-1 messages are 8-byte aligned
    assert o - (dset + 16) == u32(dset + 8)                                     # object header size
    assert sorted(seen) == [0x0001, 0x0003, 0x0005, 0x0008]
    sp = seen[0x0001]                                                           # IV.A.2.b dataspace v1: rank 1, no max dims, dim = 2
    assert (u8(sp), u8(sp + 1), u8(sp + 2)) == (1, 1, 0) and u64(sp + 8) == 2
    ty = seen[0x0003]                                                           # IV.A.2.d datatype: class 6 (compound) version 1, 16 members, 64 bytes
    assert u8(ty) == 0x16 and (u8(ty + 1) | u8(ty + 2) << 8) == 16 and u32(ty + 4) == 64
    m = ty + 8
    for i, name in enumerate(H5_NAMES):
        n = len(name) + 1
        assert b[m:m + n] == name.encode() + b"\0"
        m += (n + 7) & ~7                                                       # name padded to a multiple of 8
        assert u32(m) == 4 * i                                                  # byte offset of the member
        assert u8(m + 4) == 0                                                   # dimensionality
        m += 4 + 4 + 4 + 4 + 16                                                 # offset, dim/reserved, permutation, reserved, 4 dimension sizes
        cls, bits0, size = u8(m) & 0x0F, u8(m + 1), u32(m + 4)
        assert size == 4
        if i < 8:    # class 1 floating point, little endian, IEEE: bit offset 0, precision 32, exponent at 23 (8 bits), mantissa at 0 (23 bits), bias 127
            assert cls == 1 and bits0 & 1 == 0
            assert (u16(m + 8), u16(m + 10), u8(m + 12), u8(m + 13), u8(m + 14), u8(m + 15), u32(m + 16)) == (0, 32, 23, 8, 0, 23, 127)
            assert (u8(m + 1) >> 4) & 3 == 2 and u8(m + 2) == 31             # mantissa normalisation: implied leading 1; sign bit at 31
            m += 8 + 12
        else:        # class 0 fixed point, little endian, two's complement (signed): bit offset 0, precision 32
            assert cls == 0 and bits0 & 1 == 0 and bits0 & 8 == 8
            assert (u16(m + 8), u16(m + 10)) == (0, 32)
            m += 8 + 4
    la = seen[0x0008]                                                           # IV.A.2.i data layout v3, class 1 (contiguous): address, size
    assert (u8(la), u8(la + 1)) == (3, 1)
    data, size = u64(la + 2), u64(la + 10)
    assert size == 2 * 64 and data + size == len(b) and data % 8 == 0
    assert b[data:] == rec.tobytes()                                            # the records, verbatim (src/crixus.cpp:40-57)


def test_h5mini_rejects_corruption(tmp_path, cxlib):
    import h5mini
    rec = _records(3)
    path = str(tmp_path / "x.h5sph")
    cxlib.cx_write_h5sph_records(rec.ctypes.data, 3, path.encode())
    raw = bytearray(open(path, "rb").read())
    raw[0] ^= 0xff
    open(path, "wb").write(raw)
    with pytest.raises(h5mini.H5Error):
        h5mini.read_compound_dataset(path)
    raw[0] ^= 0xff
    open(path, "wb").write(raw[:-10])          # truncated data: end-of-file address no longer matches
    with pytest.raises(h5mini.H5Error):
        h5mini.read_compound_dataset(path)


def _stl_read(cxlib, path):
    c, n, nbe = C.POINTER(C.c_float)(), C.POINTER(C.c_float)(), api.I64(0)
    rc = cxlib.cx_stl_read(path.encode(), C.byref(c), C.byref(n), C.byref(nbe))
    out = None
    if rc == 0:
        k = int(nbe.value)
        out = (np.ctypeslib.as_array(c, shape=(k, 3, 3)).copy() if k else np.zeros((0, 3, 3), np.float32),
               np.ctypeslib.as_array(n, shape=(k, 3)).copy() if k else np.zeros((0, 3), np.float32))
    cxlib.cx_free(c)
    cxlib.cx_free(n)
    return rc, out


def test_stl_reader_edge_cases(tmp_path, cxlib):
    v, t, n = mg.plate(3, 0.1)
    good = str(tmp_path / "good.stl")
    mg.write_stl(good, v[t], n)
    rc, out = _stl_read(cxlib, good)
    assert rc == 0
    np.testing.assert_array_equal(out[0], v[t])
    np.testing.assert_array_equal(out[1], n)
    # the shipped free-surface fixture of the reference (4 triangles, VTK header)
    shipped = "/root/reference/resources/spheric2_fshape.stl"
    if os.path.exists(shipped):
        rc, out = _stl_read(cxlib, shipped)
        assert rc == 0 and out[0].shape == (4, 3, 3)
        assert abs(float(out[0][..., 2].max()) - 0.551) < 1e-6
    raw = open(good, "rb").read()
    # ASCII STL ("solid" in the first five bytes) is refused: crixus.cu:137-153 -> STL_NOT_BINARY
    asc = str(tmp_path / "ascii.stl")
    open(asc, "wb").write(b"solid x" + raw[7:])
    assert _stl_read(cxlib, asc)[0] == -1
    # record count larger than the file content: crixus.cu:218-221 -> READ_ERROR, but only when more than the last facet is
    # missing (the reference's loop still appends the facet during which it hits the end; see the fuzz test below)
    trunc = str(tmp_path / "trunc.stl")
    open(trunc, "wb").write(raw[:-130])
    assert _stl_read(cxlib, trunc)[0] == -5
    # missing file -> FILE_NOT_OPEN; empty mesh (0 facets) reads fine
    assert _stl_read(cxlib, str(tmp_path / "missing.stl"))[0] == -3
    empty = str(tmp_path / "empty.stl")
    open(empty, "wb").write(b"\0" * 80 + struct.pack("<I", 0))
    rc, out = _stl_read(cxlib, empty)
    assert rc == 0 and out[0].shape[0] == 0


def test_run_ini_without_config_or_gpu(cxlib, tmp_path, capfd):
    assert cxlib.cx_run_ini(None) == -2                               # NO_FILE, crixus.cu:57-65
    assert cxlib.cx_run_ini(str(tmp_path / "nope.ini").encode()) == -4  # CANT_READ_CONFIG, crixus.cu:67-71
    ini = _write(tmp_path, "[mesh]\nstlfile=%s\ndr=0.1\n" % (tmp_path / "missing.stl"))
    assert cxlib.cx_run_ini(ini.encode()) == -3                        # FILE_NOT_OPEN
    # input errors that the reference reports from its file readers are returned before any GPU work:
    v, t, n = mg.plate(3, 0.1)
    stl, fs, sb = str(tmp_path / "m.stl"), str(tmp_path / "m_fs.stl"), str(tmp_path / "m_sb.stl")
    mg.write_stl(stl, v[t], n)
    mg.write_stl(fs, v[t], n)
    raw = open(fs, "rb").read()
    open(fs, "wb").write(raw[:-120])                                   # free-surface file two facets short: READ_ERROR (crixus.cu:869-872)
    open(sb, "wb").write(raw[:-120])
    base = "[mesh]\nstlfile=%s\ndr=0.08\nfshape=%s\n" % (stl, fs)
    ini = _write(tmp_path, base + "[fill_0]\noption=geometry\nxseed=0.1\nyseed=0.1\nzseed=0.1\n", "g.ini")
    assert cxlib.cx_run_ini(ini.encode()) == -5
    ini = _write(tmp_path, "[mesh]\nstlfile=%s\ndr=0.08\n[special_boundary_grids]\nmesh1=%s\n" % (stl, sb), "s.ini")
    assert cxlib.cx_run_ini(ini.encode()) == -5                        # special-boundary grid short of facets (crixus.cu:564-567)
    open(stl, "wb").write(b"solid " + raw[6:])
    assert cxlib.cx_run_ini(ini.encode()) == -1                        # STL_NOT_BINARY
    capfd.readouterr()


def _host_job(nv, nbe, dimg, seed=0, sb=True):
    """A cx_job holding synthetic RESULT arrays (as cx_preprocess_host leaves them) for the CPU-only assembly tests."""
    rng = np.random.default_rng(seed)
    keep = {}
    keep["pos"] = rng.standard_normal((nv + nbe, 4)).astype(np.float32)
    keep["pos"][rng.integers(0, nv, max(nv // 10, 1)), 0] = np.float32(-1e10)        # periodic-deleted vertices
    alive = np.nonzero(keep["pos"][:nv, 0] > -1e9)[0]
    keep["norm"] = rng.standard_normal((nbe, 4)).astype(np.float32)
    ep = np.zeros((nbe, 4), dtype=np.int32)
    ep[:, :3] = alive[rng.integers(0, alive.size, (nbe, 3))]
    keep["ep"] = ep
    keep["surf"] = rng.random(nbe).astype(np.float32)
    keep["vol"] = rng.random(nv).astype(np.float32)
    G = int(np.prod(dimg))
    bits = rng.random(G) < 0.3
    keep["fpos"] = np.packbits(np.concatenate([bits, np.zeros((-G) % 32, dtype=bool)]), bitorder="little").view(np.uint32).copy()
    keep["sbid"] = rng.integers(0, 4, nv + nbe).astype(np.int32)
    job = api.CxJob()
    job.nbe, job.nvert, job.dr, job.nfluid = nbe, nv, 0.05, int(bits.sum())
    for k, t in (("pos", C.c_float), ("norm", C.c_float), ("ep", C.c_int32), ("surf", C.c_float), ("vol", C.c_float), ("fpos", C.c_uint32)):
        setattr(job, k, keep[k].ctypes.data_as(C.POINTER(t)))
    if sb:
        job.sbid = keep["sbid"].ctypes.data_as(C.POINTER(C.c_int32))
    for j in range(3):
        job.dimg[j] = dimg[j]
        job.params_fill.fcmin[j] = 0.1 * (j + 1)
    job.params_fill.dr = 0.05
    return job, keep, bits


@pytest.mark.parametrize("sb", [False, True])
def test_assemble_records_kent_grouping(sb, cxlib):
    """cx_assemble_records against a numpy restatement of crixus.cu:973-1097: fluid particles in bit order, alive vertices with
    KENT = sbid, segments stably grouped by KENT (counting sort :1088-1097), AbsoluteIndex = final position,
    VertexParticle = nfluid + v - (deleted vertices before v)."""
    nv, nbe, dimg = 120, 300, (7, 5, 4)
    job, keep, bits = _host_job(nv, nbe, dimg, seed=4, sb=sb)
    nfl = int(bits.sum())
    n = cxlib.cx_assemble_records(C.byref(job), nfl, None, 0)
    rec = np.zeros(n, dtype=REC)
    assert cxlib.cx_assemble_records(C.byref(job), nfl, rec.ctypes.data, n) == n
    alive = keep["pos"][:nv, 0] > -1e9
    assert n == nfl + int(alive.sum()) + nbe
    np.testing.assert_array_equal(rec["iref"], np.arange(n))
    np.testing.assert_array_equal(rec["kpar"], np.concatenate([np.full(nfl, 1), np.full(int(alive.sum()), 2), np.full(nbe, 3)]))
    idx = np.nonzero(bits)[0]
    dr = np.float32(0.05)
    fx = np.float32(0.1) + dr * (idx % dimg[0]).astype(np.float32)
    fz = np.float32(0.1 * 3) + dr * (idx // (dimg[0] * dimg[1])).astype(np.float32)
    np.testing.assert_array_equal(rec["x"][:nfl], fx)
    np.testing.assert_array_equal(rec["z"][:nfl], fz)
    assert (rec["vol"][:nfl] == np.float32(float(np.float32(0.05)) ** 3)).all() and not rec["kent"][:nfl].any()
    sbid = keep["sbid"] if sb else np.zeros(nv + nbe, dtype=np.int32)
    v = rec[nfl:nfl + int(alive.sum())]
    np.testing.assert_array_equal(v["kent"], sbid[:nv][alive])
    np.testing.assert_array_equal(v["vol"], keep["vol"][alive])
    order = np.argsort(sbid[nv:], kind="stable")
    s = rec[nfl + int(alive.sum()):]
    np.testing.assert_array_equal(s["kent"], sbid[nv:][order])
    np.testing.assert_array_equal(s["surf"], keep["surf"][order])
    np.testing.assert_array_equal(s["x"], keep["pos"][nv:, 0][order])
    nvs = np.cumsum(~alive)
    e = keep["ep"][order, :3]
    np.testing.assert_array_equal(np.stack([s["ep1"], s["ep2"], s["ep3"]], 1), nfl + e - nvs[e])


@pytest.mark.parametrize("fmt", [1, 2])
def test_split_output(fmt, tmp_path, cxlib, capfd, monkeypatch):
    """[output] split=true (crixus.cu:1137-1216): <name>.fluid + one file per special-boundary id holding the non-fluid
    records of that id, in order, unchanged; ids without particles are skipped."""
    from vtu import read_vtu
    import h5mini
    job, keep, bits = _host_job(90, 200, (6, 5, 3), seed=9)
    keep["sbid"][keep["sbid"] == 2] = 0           # id 2 is empty: no file
    nfl = int(bits.sum())
    n = cxlib.cx_assemble_records(C.byref(job), nfl, None, 0)
    rec = np.zeros(n, dtype=REC)
    cxlib.cx_assemble_records(C.byref(job), nfl, rec.ctypes.data, n)
    monkeypatch.chdir(tmp_path)                    # like the reference, "0." is prepended to the whole name (crixus.cpp:268)
    base = "run"
    assert cxlib.cx_write_split_records(rec.ctypes.data, n, nfl, 5, base.encode(), fmt) == 0
    out = capfd.readouterr().out
    assert " - Fluid particles: %d" % nfl in out and " - Boundary particles, kent 0: %d" % (nfl + int((rec["kent"][nfl:] == 0).sum())) in out

    def load(name):
        if fmt == 1:
            d = read_vtu(str(tmp_path / (name + ".vtu")))
            return d["AbsoluteIndex"].astype(np.int64), d["KENT"].astype(np.int64), d["Volume"], d["Surface"]
        h, info = h5mini.read_compound_dataset(str(tmp_path / ("0." + name + ".h5sph")))
        return h["AbsoluteIndex"].astype(np.int64), h["KENT"].astype(np.int64), h["Volume"], h["Surface"]
    iref, kent, vol, surf = load("run.fluid")
    np.testing.assert_array_equal(iref, np.arange(nfl))
    for k in range(5):
        want = np.nonzero(rec["kent"][nfl:] == k)[0] + nfl
        exists = os.path.exists(str(tmp_path / ("run.boundary.kent%d.vtu" % k if fmt == 1 else "0.run.boundary.kent%d.h5sph" % k)))
        assert exists == (want.size > 0), k
        if want.size:
            iref, kent, vol, surf = load("run.boundary.kent%d" % k)
            np.testing.assert_array_equal(iref, want)
            assert (kent == k).all()
            np.testing.assert_array_equal(vol, rec["vol"][want])
            np.testing.assert_array_equal(surf, rec["surf"][want])
    assert cxlib.cx_write_split_records(rec.ctypes.data, n, nfl, 2, base.encode(), fmt) == -18   # a KENT >= num_kents: INTERNAL_ERROR
    assert cxlib.cx_write_split_records(rec.ctypes.data, n, n + 1, 5, base.encode(), fmt) == -11


def test_ini_reader_fuzz_against_reference(tmp_path, cxlib, refshim):
    """600 generated .ini files built from the characters and line shapes inih treats specially (section brackets, '=' and
    ':' separators, ';' after whitespace, '#', continuation lines, over-long lines / names / sections, BOM, CR LF, empty
    names and values, repeated keys, garbage): ParseError and every lookup -- Get, GetInteger, GetReal, GetBoolean, in
    mixed case -- must equal the reference's own INIReader (vendored inih, linked into oracle/_ref/libcrixus_refshim.so)."""
    R = refshim.R
    R.refshim_ini_open.restype = C.c_void_p
    R.refshim_ini_open.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    R.refshim_ini_close.argtypes = [C.c_void_p]
    R.refshim_ini_get.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    R.refshim_ini_get_integer.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_long]
    R.refshim_ini_get_integer.restype = C.c_long
    R.refshim_ini_get_real.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_double]
    R.refshim_ini_get_real.restype = C.c_double
    R.refshim_ini_get_boolean.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    rng = np.random.default_rng(20140327)
    secs = ["mesh", "Mesh", "fill_0", "fill_1", "a b", "", "s" * 60, "x]y", "output"]
    names = ["dr", "DR", "stlfile", "x seed", "", "n" * 70, "k;v", "a=b", "use", "name"]
    vals = ["1", "0x1F", "-3.5e-2", "true", "Yes", "OFF", "maybe", "", "a ; b", "a;b", "tank.stl ;c", "v" * 250, "  padded  ", "1e400", "010",
            "+7", ".5", "nan", "0x", "2 3", "=", ":", "[x]"]

    def line():
        k = rng.integers(0, 12)
        sec, name, val = secs[rng.integers(len(secs))], names[rng.integers(len(names))], vals[rng.integers(len(vals))]
        sep = [" = ", "=", ":", " : ", "\\t=\\t"][rng.integers(5)].replace("\\t", "\t")
        if k == 0:
            return "[%s]" % sec
        if k == 1:
            return "  [%s]  ; c" % sec
        if k == 2:
            return "[%s" % sec
        if k == 3:
            return "; comment %s" % val
        if k == 4:
            return "# %s" % val
        if k == 5:
            return "   " + val               # continuation (or an error / a key, depending on what precedes it)
        if k == 6:
            return ""
        if k == 7:
            return name                      # no separator
        if k == 8:
            return name + sep + val + " ; trailing"
        return name + sep + val

    checked = 0
    for idx in range(600):
        nl = ["\n", "\r\n"][idx % 7 == 0]
        text = ("﻿" if idx % 11 == 0 else "") + nl.join(line() for _ in range(int(rng.integers(0, 14)))) + (nl if idx % 3 else "")
        path = _write(tmp_path, text, "f%d.ini" % idx).encode()
        err = C.c_int(0)
        r = R.refshim_ini_open(path, C.byref(err))
        h = api.VP()
        mine = cxlib.cx_ini_open(path, C.byref(h))
        assert mine == err.value, (idx, mine, err.value, text)
        for sec in secs + ["MESH", "nope"]:
            for name in names + ["Use", "nope"]:
                a, b = C.create_string_buffer(1024), C.create_string_buffer(1024)
                la = R.refshim_ini_get(r, sec.encode(), name.encode(), b"DEF", a, 1024)
                lb = cxlib.cx_ini_get(h, sec.encode(), name.encode(), b"DEF", b, 1024)
                assert (la, a.value) == (lb, b.value), (idx, sec, name, a.value, b.value, text)
                if a.value != b"DEF":
                    checked += 1
                    assert R.refshim_ini_get_integer(r, sec.encode(), name.encode(), -7) == cxlib.cx_ini_get_integer(h, sec.encode(), name.encode(), -7), (idx, sec, name, a.value)
                    ra, rb = R.refshim_ini_get_real(r, sec.encode(), name.encode(), -7.5), cxlib.cx_ini_get_real(h, sec.encode(), name.encode(), -7.5)
                    assert ra == rb or (ra != ra and rb != rb), (idx, sec, name, a.value, ra, rb)
                    for d in (0, 1):
                        assert R.refshim_ini_get_boolean(r, sec.encode(), name.encode(), d) == cxlib.cx_ini_get_boolean(h, sec.encode(), name.encode(), d)
        R.refshim_ini_close(r)
        cxlib.cx_ini_close(h)
    assert checked > 1000


def test_stl_reader_on_truncated_files_equals_reference_loop(tmp_path, cxlib):
    """cx_stl_read (one bulk read) against oracle/stl_reader.cpp -- the reference's ingest loop restated with the same
    istream::read calls in the same order (crixus.cu:137-221) -- on a file cut at every length around the facet records, with
    wrong facet counts, and on a few megabytes of facets (several read chunks): same return code, same facets, byte for
    byte.  Notably a file whose LAST facet is cut short is accepted by the reference: the facet is completed with the previous
    facet's values."""
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "liboracle_stl.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(["bash", os.path.join(os.path.dirname(so), "build_oracle.sh")])
    O = C.CDLL(so)
    O.orc_stl_read.argtypes = [C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_long, C.POINTER(C.c_long)]
    rng = np.random.default_rng(3)

    def both(path, cap):
        oc, on = np.zeros(cap * 9 + 9, dtype=np.float32), np.zeros(cap * 3 + 3, dtype=np.float32)
        nf = C.c_long(0)
        orc_rc = O.orc_stl_read(path.encode(), oc.ctypes.data_as(C.POINTER(C.c_float)), on.ctypes.data_as(C.POINTER(C.c_float)), cap + 1, C.byref(nf))
        rc, out = _stl_read(cxlib, path)
        assert rc == orc_rc, (path, rc, orc_rc)
        if rc == 0:
            assert out[0].shape[0] == nf.value
            assert out[0].tobytes() == oc[: 9 * nf.value].tobytes() and out[1].tobytes() == on[: 3 * nf.value].tobytes()
        return rc, (out[0].shape[0] if rc == 0 else -1)

    nfac = 7
    tris = rng.standard_normal((nfac, 3, 3)).astype(np.float32)
    nrm = rng.standard_normal((nfac, 3)).astype(np.float32)
    good = str(tmp_path / "g.stl")
    mg.write_stl(good, tris, nrm)
    raw = open(good, "rb").read()
    assert len(raw) == 84 + 50 * nfac
    accepted_short = 0
    for cut in list(range(84, len(raw) + 1)):           # every length from "count only" to the whole file
        p = str(tmp_path / "t.stl")
        open(p, "wb").write(raw[:cut])
        rc, n = both(p, nfac)
        if rc == 0 and cut < len(raw) - 2:
            accepted_short += 1
    assert accepted_short >= 48                         # the last facet may be cut anywhere (or be missing entirely)
    for count in (0, 1, nfac - 1, nfac + 1, nfac + 2, 2 ** 31 + 5, 2 ** 32 - 1):   # header count disagrees with the content
        p = str(tmp_path / "c.stl")
        open(p, "wb").write(raw[:80] + struct.pack("<I", count) + raw[84:])
        both(p, nfac + 2)
    big_n = 70000                                       # > one 65536-facet read chunk
    bt, bn = rng.standard_normal((big_n, 3, 3)).astype(np.float32), rng.standard_normal((big_n, 3)).astype(np.float32)
    big = str(tmp_path / "big.stl")
    mg.write_stl(big, bt, bn)
    assert both(big, big_n) == (0, big_n)
    rawb = open(big, "rb").read()
    open(big, "wb").write(rawb[:-17])
    assert both(big, big_n) == (0, big_n)
    open(big, "wb").write(rawb[:-67])
    assert both(big, big_n)[0] == -5
