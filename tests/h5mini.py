"""Independent minimal HDF5 reader (test infrastructure): follows the HDF5 File Format Specification far enough to
open a file the way libhdf5 would -- superblock v0 -> root symbol-table entry -> B-tree/local heap -> symbol-table
node -> dataset object header (v1) -> dataspace, compound datatype, contiguous layout -> raw data.  Written from the
specification, not from the writer in crixus_b200/csrc/host/output.cpp, so that a layout mistake there shows up here."""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"


class H5Error(ValueError):
    pass


def _u(b, off, n):
    return int.from_bytes(b[off:off + n], "little")


def _datatype(b, off):
    """-> (numpy dtype or list of members, size, bytes consumed)"""
    cv, bf0, bf1, bf2 = b[off], b[off + 1], b[off + 2], b[off + 3]
    cls, ver = cv & 0x0f, cv >> 4
    size = _u(b, off + 4, 4)
    p = off + 8
    if cls == 0:      # fixed point: bit offset, precision
        if bf0 & 1:
            raise H5Error("big-endian integer")
        signed = bool(bf0 & 0x08)
        if _u(b, p, 2) != 0 or _u(b, p + 2, 2) != 8 * size:
            raise H5Error("integer bit offset / precision")
        return np.dtype("<%s%d" % ("i" if signed else "u", size)), size, p + 4 - off
    if cls == 1:      # floating point
        if bf0 & 1:
            raise H5Error("big-endian float")
        bitoff, prec, eloc, esize, mloc, msize, ebias = struct.unpack_from("<HHBBBBI", b, p)
        if (size, bitoff, prec, eloc, esize, mloc, msize, ebias, bf1) != (4, 0, 32, 23, 8, 0, 23, 127, 31):
            raise H5Error("not IEEE binary32: %r" % ((size, bitoff, prec, eloc, esize, mloc, msize, ebias, bf1),))
        if (bf0 >> 4) & 3 != 2:
            raise H5Error("mantissa normalisation must be 'implied 1'")
        return np.dtype("<f4"), size, p + 12 - off
    if cls == 6:      # compound
        nmemb = bf0 | (bf1 << 8)
        members = []
        for _ in range(nmemb):
            end = b.index(b"\0", p)
            name = b[p:end].decode("ascii")
            if ver < 3:
                p += (end - p + 8) // 8 * 8          # name padded to a multiple of 8 (including the terminator)
            else:
                p = end + 1
            if ver == 1:
                moff = _u(b, p, 4)
                p += 4 + 1 + 3 + 4 + 4 + 16          # offset, dimensionality, reserved, permutation, reserved, 4 dims
            elif ver == 2:
                moff = _u(b, p, 4)
                p += 4
            else:
                nb = max(1, (size.bit_length() + 7) // 8)
                moff = _u(b, p, nb)
                p += nb
            mdt, msize, used = _datatype(b, p)
            p += used
            members.append((name, moff, mdt))
        return members, size, p - off
    raise H5Error("datatype class %d not supported" % cls)


def read_compound_dataset(path, name="Compound"):
    """-> (structured numpy array, dict(info)) for a contiguous 1-D compound dataset in the root group"""
    b = open(path, "rb").read()
    if b[:8] != SIG:
        raise H5Error("signature")
    if b[8] != 0:
        raise H5Error("superblock version %d" % b[8])
    so, sl = b[13], b[14]
    if (so, sl) != (8, 8):
        raise H5Error("size of offsets/lengths")
    base, eof = _u(b, 24, 8), _u(b, 40, 8)
    if base != 0 or eof != len(b):
        raise H5Error("base address %d / end of file %d vs %d bytes" % (base, eof, len(b)))
    # root group symbol table entry at 56: link name offset, object header address, cache type, reserved, scratch
    root_hdr, cache = _u(b, 64, 8), _u(b, 72, 4)
    if cache != 1:
        raise H5Error("root entry cache type")
    btree, heap = _u(b, 80, 8), _u(b, 88, 8)
    # the root object header must carry the same symbol-table message
    if b[root_hdr] != 1:
        raise H5Error("root object header version")
    nmsg, hsize = _u(b, root_hdr + 2, 2), _u(b, root_hdr + 8, 4)
    p, found = root_hdr + 16, False
    for _ in range(nmsg):
        mt, ms = _u(b, p, 2), _u(b, p + 2, 2)
        if mt == 0x11:
            found = (_u(b, p + 8, 8), _u(b, p + 16, 8)) == (btree, heap)
        p += 8 + ms
    if not found or p - (root_hdr + 16) > hsize:
        raise H5Error("root symbol-table message")
    if b[heap:heap + 4] != b"HEAP":
        raise H5Error("local heap signature")
    heap_size, heap_data = _u(b, heap + 8, 8), _u(b, heap + 24, 8)
    if b[btree:btree + 4] != b"TREE" or b[btree + 4] != 0 or b[btree + 5] != 0:
        raise H5Error("group B-tree node (leaf expected)")
    nent = _u(b, btree + 6, 2)
    dset_hdr = None
    for e in range(nent):
        child = _u(b, btree + 24 + 8 + e * 16, 8)
        if b[child:child + 4] != b"SNOD":
            raise H5Error("symbol table node signature")
        nsym = _u(b, child + 6, 2)
        for s in range(nsym):
            ent = child + 8 + s * 40
            noff = _u(b, ent, 8)
            if noff >= heap_size:
                raise H5Error("link name outside the heap")
            nm = b[heap_data + noff: b.index(b"\0", heap_data + noff)].decode("ascii")
            if nm == name:
                dset_hdr = _u(b, ent + 8, 8)
    if dset_hdr is None:
        raise H5Error("dataset %r not found" % name)
    if b[dset_hdr] != 1:
        raise H5Error("dataset object header version")
    nmsg = _u(b, dset_hdr + 2, 2)
    p = dset_hdr + 16
    dims = members = rec = addr = nbytes = None
    for _ in range(nmsg):
        mt, ms = _u(b, p, 2), _u(b, p + 2, 2)
        q = p + 8
        if mt == 0x0001:
            if b[q] != 1:
                raise H5Error("dataspace version")
            rank, flags = b[q + 1], b[q + 2]
            dims = [_u(b, q + 8 + 8 * i, 8) for i in range(rank)]
            if flags & 1:
                raise H5Error("unexpected max dims")
        elif mt == 0x0003:
            members, rec, _used = _datatype(b, q)
        elif mt == 0x0008:
            if b[q] != 3 or b[q + 1] != 1:
                raise H5Error("layout must be v3 contiguous")
            addr, nbytes = _u(b, q + 2, 8), _u(b, q + 10, 8)
        p += 8 + ms
    if None in (dims, members, rec, addr, nbytes):
        raise H5Error("dataset header incomplete")
    if len(dims) != 1 or nbytes != dims[0] * rec or addr + nbytes > len(b):
        raise H5Error("dataspace / layout mismatch")
    dt = np.dtype({"names": [m[0] for m in members], "formats": [m[2] for m in members], "offsets": [m[1] for m in members],
                   "itemsize": rec})
    return np.frombuffer(b, dtype=dt, count=dims[0], offset=addr), dict(dims=dims, record=rec, members=members, data_address=addr)
