"""Parser for the .vtu files written by the reference's vtk_output (/root/reference/src/crixus.cpp:62-259) and by our
own writer: XML header with appended-raw arrays, each prefixed by an int32 byte count."""
import re

import numpy as np

_DT = {"Float32": "<f4", "Float64": "<f8", "UInt32": "<u4", "Int32": "<i4"}


def read_vtu(path):
    raw = open(path, "rb").read()
    cut = raw.index(b"<AppendedData encoding='raw'>")
    start = raw.index(b"_", cut) + 1
    head = raw[:cut].decode("ascii", "replace")
    n = int(re.search(r"NumberOfPoints='(\d+)'", head).group(1))
    out = {"n": n}
    for m in re.finditer(r"<DataArray type='(\w+)'(?: Name='(\w+)')?(?: NumberOfComponents='(\d+)')? format='appended' offset='(\d+)'/>", head):
        typ, name, nc, off = m.group(1), m.group(2) or "Points", int(m.group(3) or 1), int(m.group(4))
        nbytes = int(np.frombuffer(raw, dtype="<i4", count=1, offset=start + off)[0])
        a = np.frombuffer(raw, dtype=_DT[typ], count=nbytes // np.dtype(_DT[typ]).itemsize, offset=start + off + 4)
        out[name] = a.reshape(n, nc) if nc > 1 else a
    return out


def split_particles(d):
    """-> dict with fluid / vertex / segment record arrays (ParticleType 1 / 2 / 3)."""
    kp = d["ParticleType"]
    res = {}
    for name, k in (("fluid", 1), ("vertex", 2), ("segment", 3)):
        m = kp == k
        res[name] = {key: val[m] for key, val in d.items() if isinstance(val, np.ndarray) and val.shape[0] == d["n"]}
    return res
