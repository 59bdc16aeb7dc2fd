"""TEST INFRASTRUCTURE -- ctypes bindings for the CPU oracle (oracle/liboracle.so, our C restatement) and for the
reference's own kernels compiled for the host (oracle/_ref/libcrixus_refshim.so), plus a small host driver that
strings the stages together the way /root/reference/src/crixus.cu:214-467,698-961 does.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


class Params(C.Structure):
    """orc_params of oracle/crixus_oracle.c == the __constant__ symbols of src/crixus_d.cu:13-32."""
    _fields_ = [("dr", C.c_float), ("dr_wall", C.c_float), ("eps", C.c_float),
                ("dmin", C.c_float * 4), ("dmax", C.c_float * 4), ("griddr", C.c_float * 4), ("gridn", C.c_int * 4),
                ("fcmin", C.c_float * 4), ("fcmax", C.c_float * 4), ("per", C.c_int * 3)]

    def copy(self):
        q = Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Params))
        return q


def _f3(a):
    return (C.c_float * 3)(*[float(x) for x in a])


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "crixus_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["bash", os.path.join(HERE, "build_oracle.sh")], stdout=subprocess.DEVNULL)
    return so


class Oracle:
    """liboracle.so -- stage functions on numpy arrays in the reference's layouts (float4 / int4 AoS)."""
    name = "oracle"

    def __init__(self):
        L = C.CDLL(build())
        P = C.POINTER(Params)
        L.orc_params_domain.argtypes = [P, C.c_float, C.c_float * 3, C.c_float * 3]
        L.orc_params_fill_eps.argtypes = [P]
        L.orc_params_container.argtypes = [P, C.c_int, C.c_float * 3, C.c_float * 3]
        L.orc_fill_dims.argtypes = [P, i64p]
        L.orc_seed_index.argtypes = [P, C.c_float * 3]
        L.orc_seed_index.restype = C.c_int64
        L.orc_weld.argtypes = [f32p, C.c_int64, C.c_float, f32p, i32p]
        L.orc_weld.restype = C.c_int64
        L.orc_swap_normals.argtypes = [f32p, C.c_int64]
        L.orc_set_bound_elem.argtypes = [f32p, f32p, f32p, i32p, C.c_int64, C.c_int64, f32p]
        L.orc_bin_segments.argtypes = [P, f32p, f32p, f32p, f32p, i32p, i32p, f32p, f32p, i32p, C.c_int64, C.c_int64]
        L.orc_find_links.argtypes = [P, f32p, C.c_int64, i32p, C.c_int]
        L.orc_find_links.restype = C.c_int64
        L.orc_periodicity_links.argtypes = [i32p, C.c_int64, i32p]
        L.orc_calc_trisize.argtypes = [i32p, i32p, C.c_int64]
        L.orc_identify_special_boundary.argtypes = [P, f32p, i32p, C.c_int64, C.c_int64, f32p, i32p, C.c_int64, i32p, C.c_int32]
        L.orc_triangle_size.argtypes = [f32p, C.c_int64, C.c_float, f32p]
        L.orc_triangle_size.restype = C.c_int64
        L.orc_calc_vert_volume.argtypes = [P, f32p, f32p, i32p, f32p, i32p, i32p, C.c_int64, C.c_int64, C.c_int]
        L.orc_calc_vert_volume.restype = C.c_int64
        L.orc_fill_box.argtypes = [P, u32p] + [C.c_float] * 6
        L.orc_fill_box.restype = C.c_int64
        L.orc_fill_geometry.argtypes = [P, u32p, f32p, i32p, f32p, C.c_int64, C.c_int64, i32p, C.c_int64, C.POINTER(C.c_int)]
        L.orc_fill_geometry.restype = C.c_int64
        L.orc_fill_geometry_slab.argtypes = [P, u32p, f32p, i32p, f32p, C.c_int64, C.c_int64, i32p, C.c_int64, C.c_int64, C.c_int64,
                                             C.POINTER(C.c_int)]
        L.orc_fill_geometry_slab.restype = C.c_int64
        L.orc_check_collision.argtypes = [P, i64p, i64p, f32p, i32p, f32p, C.c_int64, C.c_int64, i32p]
        L.orc_num_threads.restype = C.c_int
        self.L = L

    # ---- params
    def params_domain(self, dr, bbmin, bbmax):
        p = Params()
        self.L.orc_params_domain(C.byref(p), dr, _f3(bbmin), _f3(bbmax))
        return p

    def params_fill_eps(self, p):
        self.L.orc_params_fill_eps(C.byref(p))

    def params_container(self, p, use, cmin=(0, 0, 0), cmax=(0, 0, 0)):
        self.L.orc_params_container(C.byref(p), int(use), _f3(cmin), _f3(cmax))

    def fill_dims(self, p):
        d = np.zeros(3, dtype=np.int64)
        self.L.orc_fill_dims(C.byref(p), d)
        return d

    def seed_index(self, p, spos):
        return int(self.L.orc_seed_index(C.byref(p), _f3(spos)))

    # ---- stages
    def weld(self, corners, dr):
        corners = np.ascontiguousarray(corners, dtype=np.float32).reshape(-1, 3)
        out_v = np.zeros_like(corners)
        out_id = np.zeros(corners.shape[0], dtype=np.int32)
        nv = self.L.orc_weld(corners, corners.shape[0], dr, out_v, out_id)
        return out_v[:nv].copy(), out_id.reshape(-1, 3)

    def swap_normals(self, p, norm):
        self.L.orc_swap_normals(norm, norm.shape[0])

    def set_bound_elem(self, p, pos, norm, surf, ep, nbe, nvert):
        z = np.zeros(4, dtype=np.float32)
        self.L.orc_set_bound_elem(pos, norm, surf, ep, nbe, nvert, z)
        return z

    def bin_segments(self, p, pre_pos, pre_norm, pre_ep, pre_surf, nbe, nvert):
        pos = np.zeros_like(pre_pos)
        norm = np.zeros_like(pre_norm)
        ep = np.zeros_like(pre_ep)
        surf = np.zeros_like(pre_surf)
        cell_idx = np.zeros(p.gridn[3] + 1, dtype=np.int32)
        self.L.orc_bin_segments(C.byref(p), pre_pos, pos, pre_norm, norm, pre_ep, ep, pre_surf, surf, cell_idx, nbe, nvert)
        return pos, norm, ep, surf, cell_idx

    def find_links(self, p, pos, nvert, newlink, idim):
        return int(self.L.orc_find_links(C.byref(p), pos, nvert, newlink, idim))

    def periodicity_links(self, p, pos, ep, nvert, nbe, newlink, idim):
        self.L.orc_periodicity_links(ep, nbe, newlink)

    def identify_special_boundary(self, p, pos, ep, nvert, nbe, sbpos, sbep, sbid, sbi, cell_idx=None):
        """identifySpecialBoundarySegments (crixus_d.cu:1085-1112); sbid int32[nvert+nbe] is updated in place"""
        self.L.orc_identify_special_boundary(C.byref(p), pos, ep, nvert, nbe, sbpos, sbep, sbep.shape[0], sbid, sbi)

    def triangle_size(self, corners, dr):
        """resources/test-triangle-size.c -> (failures, min distance, max distance)"""
        corners = np.ascontiguousarray(corners, dtype=np.float32).reshape(-1)
        out = np.zeros(2, dtype=np.float32)
        n = int(self.L.orc_triangle_size(corners, corners.size // 9, np.float32(dr), out))
        return n, float(out[0]), float(out[1])

    def calc_trisize(self, p, ep, nvert, nbe):
        t = np.zeros(nvert, dtype=np.int32)
        self.L.orc_calc_trisize(ep, t, nbe)
        return t

    def calc_vert_volume(self, p, pos, norm, ep, trisize, cell_idx, nvert, nbe, zero_open=False, v0=0, v1=None):
        vol = np.zeros(nvert, dtype=np.float32)
        bad = self.L.orc_calc_vert_volume(C.byref(p), pos, norm, ep, vol, trisize, cell_idx, v0, nvert if v1 is None else v1,
                                          int(zero_open))
        if bad:
            raise RuntimeError("trisize > trimax for %d vertices" % bad)
        return vol

    def fill_box(self, p, fpos, lo, hi):
        return int(self.L.orc_fill_box(C.byref(p), fpos, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]))

    def fill_geometry(self, p, fpos, norm, ep, pos, nbe, cnbe, cell_idx, seed):
        sw = C.c_int(0)
        n = int(self.L.orc_fill_geometry(C.byref(p), fpos, norm, ep, pos, nbe, cnbe, cell_idx, seed, C.byref(sw)))
        self.last_sweeps = sw.value
        return n

    def fill_geometry_slab(self, p, fpos, norm, ep, pos, nbe, cnbe, cell_idx, seed, k0, k1, first_call=True):
        sw = C.c_int(0)
        return int(self.L.orc_fill_geometry_slab(C.byref(p), fpos, norm, ep, pos, nbe, cnbe, cell_idx, seed if first_call else -1,
                                                 k0, k1, C.byref(sw)))

    def check_collision(self, p, s, e, norm, ep, pos, nbe, cnbe, cell_idx):
        return int(self.L.orc_check_collision(C.byref(p), np.asarray(s, dtype=np.int64), np.asarray(e, dtype=np.int64),
                                              norm, ep, pos, nbe, cnbe, cell_idx))

    def num_threads(self):
        return int(self.L.orc_num_threads())


class RefShim(Oracle):
    """oracle/_ref/libcrixus_refshim.so: the reference's own src/crixus_d.cu compiled for the host (32-bit limits,
    g++ FMA contraction).  Params/geometry glue is inherited from Oracle (it is host code in the reference too)."""
    name = "refshim"

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libcrixus_refshim.so"))

    def __init__(self, threads=0):
        super().__init__()
        R = C.CDLL(os.path.join(HERE, "_ref", "libcrixus_refshim.so"))
        R.refshim_set_consts.argtypes = [C.c_float * 4, C.c_int * 4, C.c_float, C.c_float * 4, C.c_float * 4, C.c_float * 4,
                                         C.c_float * 4, C.c_int * 3, C.c_float, C.c_float]
        R.refshim_swap_normals.argtypes = [f32p, C.c_int]
        R.refshim_set_bound_elem.argtypes = [f32p, f32p, f32p, i32p, C.c_uint, C.c_int, f32p]
        R.refshim_bin.argtypes = [f32p, f32p, f32p, f32p, i32p, i32p, f32p, f32p, i32p, C.c_uint, C.c_uint]
        R.refshim_find_links.argtypes = [f32p, C.c_int, i32p, C.c_int]
        R.refshim_periodicity_links.argtypes = [f32p, i32p, C.c_int, C.c_int, i32p, C.c_int]
        R.refshim_calc_trisize.argtypes = [i32p, i32p, C.c_int]
        R.refshim_identify_sb.argtypes = [f32p, i32p, C.c_int, C.c_int, f32p, i32p, C.c_int, i32p, C.c_int]
        R.refshim_calc_vert_volume.argtypes = [f32p, f32p, i32p, f32p, i32p, i32p, C.c_int, C.c_int, C.c_int]
        R.refshim_fill_fluid.argtypes = [u32p] + [C.c_float] * 6
        R.refshim_fill_fluid.restype = C.c_uint
        R.refshim_fill_fluid_complex_sweep.argtypes = [u32p, f32p, i32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
        R.refshim_fill_fluid_complex_sweep.restype = C.c_uint
        R.refshim_weld.argtypes = [f32p, C.c_int, C.c_float, f32p, u32p]
        R.refshim_weld.restype = C.c_int
        R.refshim_set_threads.argtypes = [C.c_int]
        R.refshim_set_threads(threads)
        self.R = R

    def _consts(self, p):
        self.R.refshim_set_consts(p.griddr, p.gridn, p.eps, p.dmin, p.dmax, p.fcmin, p.fcmax, p.per, p.dr, p.dr_wall)

    def weld(self, corners, dr):
        corners = np.ascontiguousarray(corners, dtype=np.float32).reshape(-1, 3)
        out_v = np.zeros_like(corners)
        out_id = np.zeros(corners.shape[0], dtype=np.uint32)
        nv = self.R.refshim_weld(corners, corners.shape[0] // 3, dr, out_v, out_id)
        return out_v[:nv].copy(), out_id.astype(np.int32).reshape(-1, 3)

    def swap_normals(self, p, norm):
        self.R.refshim_swap_normals(norm, norm.shape[0])

    def set_bound_elem(self, p, pos, norm, surf, ep, nbe, nvert):
        self._consts(p)
        z = np.zeros(4, dtype=np.float32)
        self.R.refshim_set_bound_elem(pos, norm, surf, ep, nbe, nvert, z)
        return z

    def bin_segments(self, p, pre_pos, pre_norm, pre_ep, pre_surf, nbe, nvert):
        self._consts(p)
        pos = np.zeros_like(pre_pos)
        norm = np.zeros_like(pre_norm)
        ep = np.zeros_like(pre_ep)
        surf = np.zeros_like(pre_surf)
        cell_idx = np.zeros(p.gridn[3] + 1, dtype=np.int32)
        self.R.refshim_bin(pre_pos, pos, pre_norm, norm, pre_ep, ep, pre_surf, surf, cell_idx, nbe, nvert)
        return pos, norm, ep, surf, cell_idx

    def find_links(self, p, pos, nvert, newlink, idim):
        self._consts(p)
        self.R.refshim_find_links(pos, nvert, newlink, idim)
        return 0

    def periodicity_links(self, p, pos, ep, nvert, nbe, newlink, idim):
        self.R.refshim_periodicity_links(pos, ep, nvert, nbe, newlink, idim)

    def identify_special_boundary(self, p, pos, ep, nvert, nbe, sbpos, sbep, sbid, sbi, cell_idx=None):
        self._consts(p)
        self.R.refshim_identify_sb(pos, ep, nvert, nbe, sbpos, sbep, sbep.shape[0], sbid, sbi)

    def calc_trisize(self, p, ep, nvert, nbe):
        t = np.zeros(nvert, dtype=np.int32)
        self.R.refshim_calc_trisize(ep, t, nbe)
        return t

    def calc_vert_volume(self, p, pos, norm, ep, trisize, cell_idx, nvert, nbe, zero_open=False, v0=0, v1=None):
        self._consts(p)
        vol = np.zeros(nvert, dtype=np.float32)
        self.R.refshim_calc_vert_volume(pos, norm, ep, vol, trisize, cell_idx, nvert, nbe, int(zero_open))
        return vol

    def fill_box(self, p, fpos, lo, hi):
        self._consts(p)
        return int(self.R.refshim_fill_fluid(fpos, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]))

    def fill_geometry(self, p, fpos, norm, ep, pos, nbe, cnbe, cell_idx, seed):
        self._consts(p)
        total, it = 0, 0
        while True:  # src/crixus.cu:934-945
            it += 1
            nfi = int(self.R.refshim_fill_fluid_complex_sweep(fpos, norm, ep, pos, nbe, seed, cnbe, it, cell_idx))
            total += nfi
            if nfi == 0 or it >= 10000:
                break
        self.last_sweeps = it
        return total


# ---------------------------------------------------------------------------------------------- host driver
def f4(a3, n=None):
    """(n,3) float32 -> (n,4) float4 AoS with .w = 0."""
    a3 = np.asarray(a3, dtype=np.float32).reshape(-1, 3)
    out = np.zeros((a3.shape[0] if n is None else n, 4), dtype=np.float32)
    out[:a3.shape[0], :3] = a3
    return out


def run_pipeline(B, tris, normals, dr, swap_normals=False, periodic=(False, False, False), zero_open=False,
                 container=None, fills=(), fshape=None, stages=("volume", "fill"), special=()):
    """The hot path of crixus_main (/root/reference/src/crixus.cu:184-961) on backend B (Oracle or RefShim).

    tris (n,3,3) / normals (n,3): the STL content.  fills: list of ("box", lo, hi) or ("geometry", seed, dr_wall).
    fshape: optional (tris, normals) of the free-surface STL.  special: list of (n,3,3) corner arrays, the special-boundary
    grids #1, #2, ... (crixus.cu:469-617).  Returns a dict of all stage outputs."""
    tris = np.ascontiguousarray(tris, dtype=np.float32)
    nbe = tris.shape[0]
    corners = tris.reshape(-1, 3)
    bbmin, bbmax = corners.min(0), corners.max(0)                  # crixus.cu:185-208
    verts, epv = B.weld(corners, np.float32(dr))                   # :214
    nvert = verts.shape[0]
    p = B.params_domain(np.float32(dr), bbmin, bbmax)              # :262-283
    pre_pos = f4(verts, nvert + nbe)
    pre_norm = f4(normals)
    pre_ep = np.zeros((nbe, 4), dtype=np.int32)
    pre_ep[:, :3] = epv
    pre_surf = np.zeros(nbe, dtype=np.float32)
    if swap_normals:
        B.swap_normals(p, pre_norm)                                # :310-317
    zstat = B.set_bound_elem(p, pre_pos, pre_norm, pre_surf, pre_ep, nbe, nvert)  # :338
    pos, norm, ep, surf, cell_idx = B.bin_segments(p, pre_pos, pre_norm, pre_ep, pre_surf, nbe, nvert)  # :355-361
    out = dict(params_domain=p.copy(), nvert=nvert, nbe=nbe, verts=verts, epv=epv, zstat=zstat, pre_ep=pre_ep.copy(),
               pre_surf=pre_surf, pre_bary=pre_pos[nvert:].copy())
    missing = 0
    for idim in range(3):                                          # :391-412
        if periodic[idim]:
            newlink = np.full(nvert, -1, dtype=np.int32)
            missing += B.find_links(p, pos, nvert, newlink, idim)
            B.periodicity_links(p, pos, ep, nvert, nbe, newlink, idim)
            out["newlink%d" % idim] = newlink
    for j in range(3):
        p.per[j] = int(periodic[j])                                # :428
    out.update(pos=pos, norm=norm, ep=ep, surf=surf, cell_idx=cell_idx, missing_partners=missing)
    trisize = B.calc_trisize(p, ep, nvert, nbe)                    # :435
    out["trisize"] = trisize
    if "volume" in stages:
        out["vol"] = B.calc_vert_volume(p, pos, norm, ep, trisize, cell_idx, nvert, nbe, zero_open)  # :437
    out["params_volume"] = p.copy()
    if "fill" not in stages and not len(special):
        return out
    B.params_fill_eps(p)                                           # :464-467
    if len(special):                                               # :469-617
        sbid = np.zeros(nvert + nbe, dtype=np.int32)
        for k, sb_t in enumerate(special):
            sb_v, sb_e = B.weld(np.ascontiguousarray(sb_t, dtype=np.float32).reshape(-1, 3), np.float32(dr))   # :560
            sb_ep = np.zeros((sb_e.shape[0], 4), dtype=np.int32)
            sb_ep[:, :3] = sb_e
            B.identify_special_boundary(p, pos, ep, nvert, nbe, f4(sb_v), sb_ep, sbid, k + 1, cell_idx=cell_idx)   # :606
        out["sbid"] = sbid
    if "fill" not in stages:
        out["params_fill"] = p.copy()
        return out
    if container is not None:                                      # :698-722
        B.params_container(p, True, container[0], container[1])
    else:
        B.params_container(p, False)
    dimg = B.fill_dims(p)
    G = int(dimg[0] * dimg[1] * dimg[2])
    fpos = np.zeros((G + 31) // 32, dtype=np.uint32)
    fnorm, fep, fposa, fnbe, cnbe = norm, ep, pos[:nvert].copy(), nbe, 0
    if fshape is not None:                                         # :809-930 (first geometry fill)
        fs_t, fs_n = fshape
        fs_v, fs_e = B.weld(np.ascontiguousarray(fs_t, dtype=np.float32).reshape(-1, 3), np.float32(dr))
        cnbe = fs_e.shape[0]
        fnbe = nbe + cnbe
        fposa = np.concatenate([pos[:nvert], f4(fs_v)])
        fnorm = np.concatenate([norm, f4(fs_n)])
        fe4 = np.zeros((cnbe, 4), dtype=np.int32)
        fe4[:, :3] = fs_e + nvert
        fep = np.concatenate([ep, fe4])
    fposa, fnorm, fep = map(np.ascontiguousarray, (fposa, fnorm, fep))
    nfluid = 0
    counts = []
    for f in fills:                                                # :734-961
        if f[0] == "box":
            n = B.fill_box(p, fpos, [np.float32(x) for x in f[1]], [np.float32(x) for x in f[2]])
        else:
            p.dr_wall = np.float32(f[2] if len(f) > 2 and f[2] is not None else dr)   # :795-797
            seed = B.seed_index(p, f[1])
            n = B.fill_geometry(p, fpos, fnorm, fep, fposa, fnbe, cnbe, cell_idx, seed)
        counts.append(n)
        nfluid += n
    out.update(params_fill=p.copy(), dimg=dimg, fpos=fpos, nfluid=nfluid, fill_counts=counts, fposa=fposa, fnorm=fnorm, fep=fep,
               fnbe=fnbe, cnbe=cnbe)
    return out
