"""Comparison of a pipeline result (dict of oracle.orc.run_pipeline) with a golden fixture produced by the reference CUDA
binary (tests/golden/make_golden.py): vertex volumes in vertex order, segment records in canonical order, fluid
particle positions in bit order, particle counts."""
import numpy as np


def assemble(g):
    """particle lists the way /root/reference/src/crixus.cu:991-1097 builds them (without the 64-byte records)."""
    nv, nbe = g["nvert"], g["nbe"]
    p = g["params_fill"] if "params_fill" in g else None
    alive = g["pos"][:nv, 0] > -1e9
    shift = np.cumsum(~alive) - (~alive)          # nvshift
    nvs = np.cumsum(~alive)
    out = dict(vert_pos=g["pos"][:nv, :3][alive], vert_vol=g["vol"][alive])
    nfl = g.get("nfluid", 0)
    ep = g["ep"][:, :3].astype(np.int64)
    out["seg_ep"] = nfl + ep - nvs[ep]            # crixus.cu:1082-1084 (deleted vertices never referenced)
    out["seg_pos"] = g["pos"][nv:, :3]
    out["seg_norm"] = g["norm"][:, :3]
    out["seg_surf"] = g["surf"]
    sbid = g["sbid"] if "sbid" in g else np.zeros(nv + nbe, dtype=np.int32)
    out["vert_kent"] = sbid[:nv][alive]           # crixus.cu:1042-1045
    out["seg_kent"] = sbid[nv:]                   # crixus.cu:1072-1077
    if p is not None:
        d = g["dimg"]
        bits = np.unpackbits(g["fpos"].view(np.uint8), bitorder="little")[: int(d[0] * d[1] * d[2])]
        idx = np.nonzero(bits)[0]
        k, rem = idx // (d[0] * d[1]), idx % (d[0] * d[1])
        j, i = rem // d[0], rem % d[0]
        dr = np.float32(p.dr)
        fc = np.array(list(p.fcmin)[:3], dtype=np.float32)
        out["fluid_pos"] = np.stack([fc[0] + dr * i.astype(np.float32), fc[1] + dr * j.astype(np.float32),
                                     fc[2] + dr * k.astype(np.float32)], 1).astype(np.float32)   # crixus.cu:996-1002
    else:
        out["fluid_pos"] = np.zeros((0, 3), dtype=np.float32)
    del shift
    return out


def canonical(a):
    key = np.lexsort((a["seg_pos"][:, 2], a["seg_pos"][:, 1], a["seg_pos"][:, 0], a["seg_ep"][:, 2], a["seg_ep"][:, 1], a["seg_ep"][:, 0]))
    return {k: (v[key] if k.startswith("seg_") else v) for k, v in a.items()}


def compare_with_golden(g, gold, c, vol_exact=True):
    a = canonical(assemble(g))
    assert a["vert_pos"].shape == gold["vert_pos"].shape
    np.testing.assert_array_equal(a["vert_pos"], gold["vert_pos"])
    np.testing.assert_array_equal(a["fluid_pos"], gold["fluid_pos"])
    # The reference's running particle count `nfluid` (crixus.cu:944) can exceed the number of fluid particles it
    # writes by one: fill_fluid_complex keeps its bitfield word in a register (SASS: one LDG before the bit loop, STG
    # after a fill), so the thread that owns the seed's word can overwrite the seed set by thread 0 and then fill
    # and count the seed cell a second time (a data race; it shows on the GPU whenever the seed's word is not owned
    # by warp 0 of block 0, never in the host build).  The extra count shifts every VertexParticle index by one and
    # leaves one uninitialised record at the end of the file.  We compare modulo that shift and require it to be 0/1.
    import re
    counted = int(re.search(r"Creation of (\d+) fluid particles", str(gold["log"])).group(1))   # crixus.cu:963
    shift = counted - gold["fluid_pos"].shape[0]
    assert shift in (0, 1), shift
    np.testing.assert_array_equal(a["seg_ep"] + shift, gold["seg_ep"])
    np.testing.assert_array_equal(a["seg_pos"], gold["seg_pos"])
    np.testing.assert_array_equal(a["seg_norm"], gold["seg_norm"])
    np.testing.assert_array_equal(a["seg_surf"], gold["seg_surf"])
    if "seg_kent" in gold.files:                  # special-boundary ids (fixtures made before that stage existed have none)
        np.testing.assert_array_equal(a["vert_kent"], gold["vert_kent"])
        np.testing.assert_array_equal(a["seg_kent"], gold["seg_kent"])
        fo = gold["seg_kent_file_order"]          # the file holds the segments grouped by id (crixus.cu:1088-1097)
        assert (np.diff(fo) >= 0).all()
    else:
        assert not a["seg_kent"].any() and not a["vert_kent"].any()
    gv, ov = a["vert_vol"], gold["vert_vol"]
    q = float(np.float32(c["dr"]) / 40) ** 3
    nbad = int((gv != ov).sum())
    maxq = float(np.abs(gv.astype(np.float64) - ov.astype(np.float64)).max() / q) if gv.size else 0.0
    if vol_exact:
        assert nbad == 0, "%d of %d vertex volumes differ from the reference, max %.3f voxel quanta" % (nbad, gv.size, maxq)
    return nbad, maxq
