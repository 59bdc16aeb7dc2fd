"""ctypes binding of the C-ABI declared in include/crixus_b200.h.

PyTorch is used for device memory and streams only; every computation goes through libcrixus_b200.so (hand-written
sm_100a kernels).  There is no CPU fallback: if the library is missing or no GPU is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CX_LIB_PATH") or os.path.join(HERE, "libcrixus_b200.so")   # CX_LIB_PATH: kernel experiments (tools/)


class CxParams(C.Structure):
    """cx_params == the __constant__ symbols of /root/reference/src/crixus_d.cu:13-32."""
    _fields_ = [("dr", C.c_float), ("dr_wall", C.c_float), ("eps", C.c_float),
                ("dmin", C.c_float * 4), ("dmax", C.c_float * 4), ("griddr", C.c_float * 4), ("gridn", C.c_int32 * 4),
                ("fcmin", C.c_float * 4), ("fcmax", C.c_float * 4), ("per", C.c_int32 * 3)]

    def copy(self):
        q = CxParams()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(CxParams))
        return q


class CxFillSpec(C.Structure):
    _fields_ = [("option", C.c_int32), ("lo", C.c_float * 3), ("hi", C.c_float * 3), ("seed", C.c_float * 3),
                ("dr_wall", C.c_float)]


class CxJob(C.Structure):
    _fields_ = [("corners", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)), ("nbe", C.c_int64),
                ("welded_pos", C.POINTER(C.c_float)), ("welded_ep", C.POINTER(C.c_int32)), ("welded_nvert", C.c_int64),
                ("bbmin", C.c_float * 3), ("bbmax", C.c_float * 3), ("dr", C.c_float),
                ("swap_normals", C.c_int32), ("zero_open", C.c_int32), ("periodic", C.c_int32 * 3),
                ("use_container", C.c_int32), ("cmin", C.c_float * 3), ("cmax", C.c_float * 3),
                ("fs_corners", C.POINTER(C.c_float)), ("fs_normals", C.POINTER(C.c_float)), ("fs_nbe", C.c_int64),
                ("fills", C.POINTER(CxFillSpec)), ("nfills", C.c_int32),
                ("sb_corners", C.POINTER(C.POINTER(C.c_float))), ("sb_nbe", C.POINTER(C.c_int64)), ("nsb", C.c_int32),
                ("nvert", C.c_int64), ("nfluid", C.c_int64), ("nfluid_counted", C.c_int64),
                ("pos", C.POINTER(C.c_float)), ("norm", C.POINTER(C.c_float)), ("ep", C.POINTER(C.c_int32)),
                ("surf", C.POINTER(C.c_float)), ("vol", C.POINTER(C.c_float)), ("cell_idx", C.POINTER(C.c_int32)),
                ("fpos", C.POINTER(C.c_uint32)), ("sbid", C.POINTER(C.c_int32)),
                ("d_pos", C.c_void_p), ("d_norm", C.c_void_p), ("d_ep", C.c_void_p), ("d_surf", C.c_void_p), ("d_vol", C.c_void_p),
                ("d_fpos", C.c_void_p), ("d_sbid", C.c_void_p), ("dimg", C.c_int64 * 3), ("zstat", C.c_float * 4),
                ("params_fill", CxParams),
                ("tri_failed", C.c_int64), ("tri_mind", C.c_float), ("tri_maxd", C.c_float),
                ("t_weld_s", C.c_double), ("t_gpu_s", C.c_double), ("t_h2d_s", C.c_double), ("t_d2h_s", C.c_double),
                ("stage_ms", C.c_float * 8),
                ("out_sharded", C.c_int32),
                ("seg_lo", C.c_int64), ("seg_hi", C.c_int64), ("vert_lo", C.c_int64), ("vert_hi", C.c_int64),
                ("fword_lo", C.c_int64), ("fword_hi", C.c_int64)]


ERRORS = {0: "CX_OK", -1: "STL_NOT_BINARY", -2: "NO_FILE", -3: "FILE_NOT_OPEN", -4: "CANT_READ_CONFIG", -5: "READ_ERROR",
          -6: "FLUID_NDEF", -7: "WRITE_FAIL", -9: "NO_PER_VERT", -11: "WRONG_INPUT", -16: "NO_WRITE_FILE",
          -17: "FLUID_BBOX_TOO_SMALL", -18: "INTERNAL_ERROR", -101: "CUDA_ERROR", -102: "TRIMAX_EXCEEDED",
          -103: "FAN_INCOMPLETE", -104: "SEED_OUTSIDE", -105: "NO_DEVICE", -106: "WELD_CELL_TOO_FULL", -107: "NO_NCCL"}


class CxError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__("crixus_b200: %s (%d) %s" % (ERRORS.get(code, "?"), code, msg))
        self.code = code


_lib = None
VP = C.c_void_p
I64 = C.c_int64


def lib():
    """Loads libcrixus_b200.so (raises if it has not been built -- there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libcrixus_b200.so not built: run `python -m crixus_b200.build` (or __graft_entry__.build())")
    L = C.CDLL(LIB_PATH)
    P = C.POINTER(CxParams)
    F3 = C.c_float * 3
    sig = {
        "cx_params_domain": (None, [P, C.c_float, F3, F3]),
        "cx_params_fill_eps": (None, [P]),
        "cx_params_container": (None, [P, C.c_int, F3, F3]),
        "cx_fill_dims": (None, [P, C.POINTER(I64)]),
        "cx_seed_index": (I64, [P, F3]),
        "cx_create": (C.c_int, [C.c_int, C.POINTER(VP)]),
        "cx_destroy": (None, [VP]),
        "cx_set_stream": (C.c_int, [VP, VP]),
        "cx_synchronize": (C.c_int, [VP]),
        "cx_get_stream": (VP, [VP]),
        "cx_last_error": (C.c_char_p, [VP]),
        "cx_version": (C.c_char_p, []),
        "cx_kernel_launches": (I64, [VP]),
        "cx_measure_fp32_peak": (C.c_int, [VP, C.c_int, C.POINTER(C.c_double)]),
        "cx_swap_normals": (C.c_int, [VP, VP, I64]),
        "cx_set_bound_elem": (C.c_int, [VP, VP, VP, VP, VP, I64, I64, C.POINTER(C.c_float)]),
        "cx_bin_segments": (C.c_int, [VP, P, VP, VP, VP, VP, VP, VP, VP, VP, VP, I64, I64]),
        "cx_find_links": (C.c_int, [VP, P, VP, I64, VP, C.c_int]),
        "cx_periodicity_links": (C.c_int, [VP, VP, I64, VP]),
        "cx_calc_trisize": (C.c_int, [VP, VP, VP, I64]),
        "cx_calc_vert_volume": (C.c_int, [VP, P, VP, VP, VP, VP, VP, VP, I64, I64, C.c_int, I64, I64]),
        "cx_calc_vert_volume_part": (C.c_int, [VP, P, VP, VP, VP, VP, VP, VP, I64, I64, C.c_int, I64, C.c_int, C.c_int]),
        "cx_identify_special_boundary": (C.c_int, [VP, P, VP, VP, VP, I64, I64, VP, VP, I64, VP, C.c_int32]),
        "cx_fill_box": (C.c_int, [VP, P, VP, F3, F3, C.POINTER(I64)]),
        "cx_fill_geometry": (C.c_int, [VP, P, VP, VP, VP, VP, I64, I64, VP, I64, C.POINTER(I64), C.POINTER(C.c_int32)]),
        "cx_fill_geometry_slab": (C.c_int, [VP, P, VP, VP, VP, VP, I64, I64, VP, I64, I64, I64, C.c_int, C.POINTER(I64),
                                            C.POINTER(C.c_int32)]),
        "cx_fill_set_max_rounds": (C.c_int, [VP, C.c_int]),
        "cx_fill_unfinished": (C.c_int, [VP]),
        "cx_fill_set_resolve_mode": (C.c_int, [VP, C.c_int]),
        "cx_fill_resolve_slab": (C.c_int, [VP, P, VP, VP, VP, VP, I64, I64, VP, I64, I64, C.POINTER(VP), C.POINTER(I64)]),
        "cx_fill_resolve_part": (C.c_int, [VP, P, VP, VP, VP, VP, I64, I64, VP, C.c_int, C.c_int, C.POINTER(VP), C.POINTER(I64)]),
        "cx_fill_propagate": (C.c_int, [VP, P, VP, VP, VP, VP, I64, I64, VP, I64, C.POINTER(I64), C.POINTER(C.c_int32)]),
        "cx_preprocess_host": (C.c_int, [VP, C.POINTER(CxJob)]),
        "cx_preprocess_dist": (C.c_int, [VP, C.POINTER(CxJob)]),
        "cx_dist_unique_id": (C.c_int, [VP]),
        "cx_dist_init": (C.c_int, [VP, VP, C.c_int, C.c_int]),
        "cx_dist_attach": (C.c_int, [VP, VP, C.c_int, C.c_int]),
        "cx_dist_init_local": (C.c_int, [C.POINTER(VP), C.c_int]),
        "cx_dist_finalize": (C.c_int, [VP]),
        "cx_dist_world": (C.c_int, [VP]),
        "cx_dist_rank": (C.c_int, [VP]),
        "cx_dist_calc_vert_volume": (C.c_int, [VP, P, VP, VP, VP, VP, VP, VP, I64, I64, C.c_int]),
        "cx_dist_fill_geometry": (C.c_int, [VP, P, VP, VP, VP, VP, I64, I64, VP, I64, C.POINTER(I64), C.POINTER(C.c_int32)]),
        "cx_hash_u32": (C.c_int, [VP, VP, I64, C.POINTER(C.c_uint64)]),
        "cx_dist_set_fill_mode": (C.c_int, [VP, C.c_int]),
        "cx_job_free": (None, [C.POINTER(CxJob)]),
        "cx_job_sizeof": (I64, []),
        "cx_weld_host": (I64, [C.POINTER(C.c_float), I64, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
        "cx_bbox_device": (C.c_int, [VP, VP, I64, F3, F3]),
        "cx_check_triangle_size": (C.c_int, [VP, VP, I64, C.c_float, C.POINTER(I64), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "cx_weld_device": (C.c_int, [VP, VP, I64, C.c_float, F3, F3, VP, VP, C.POINTER(I64)]),
        "cx_ini_open": (C.c_int, [C.c_char_p, C.POINTER(VP)]),
        "cx_ini_close": (None, [VP]),
        "cx_ini_get": (C.c_int, [VP, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
        "cx_ini_get_integer": (C.c_long, [VP, C.c_char_p, C.c_char_p, C.c_long]),
        "cx_ini_get_real": (C.c_double, [VP, C.c_char_p, C.c_char_p, C.c_double]),
        "cx_ini_get_boolean": (C.c_int, [VP, C.c_char_p, C.c_char_p, C.c_int]),
        "cx_stl_read": (C.c_int, [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.POINTER(C.c_float)), C.POINTER(I64)]),
        "cx_free": (None, [VP]),
        "cx_assemble_records": (I64, [C.POINTER(CxJob), I64, VP, I64]),
        "cx_write_vtu_records": (C.c_int, [VP, I64, C.c_char_p]),
        "cx_write_h5sph_records": (C.c_int, [VP, I64, C.c_char_p]),
        "cx_write_split_records": (C.c_int, [VP, I64, I64, C.c_int32, C.c_char_p, C.c_int]),
        "cx_assemble_records_device": (C.c_int, [VP, P, C.POINTER(I64), VP, VP, VP, VP, VP, VP, VP, I64, I64, I64, VP, I64, C.POINTER(I64)]),
        "cx_write_vtu_records_device": (C.c_int, [VP, VP, I64, C.c_char_p]),
        "cx_write_h5sph_records_device": (C.c_int, [VP, VP, I64, C.c_char_p]),
        "cx_stream_h5sph_device": (C.c_int, [VP, P, C.POINTER(I64), VP, VP, VP, VP, VP, VP, I64, I64, I64, C.c_char_p, I64, C.POINTER(I64)]),
        "cx_run_ini": (C.c_int, [C.c_char_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)   # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    L._declared = sorted(sig)
    if L.cx_job_sizeof() != C.sizeof(CxJob):
        raise ImportError("crixus_b200.api.CxJob (%d bytes) does not mirror the library's cx_job (%d bytes)" % (C.sizeof(CxJob), L.cx_job_sizeof()))
    _lib = L
    return L


def F3c():
    return (C.c_float * 3)()


def f3(a):
    return (C.c_float * 3)(*[float(x) for x in a])


# ---- host-side glue (no GPU needed)
def params_domain(dr, bbmin, bbmax):
    p = CxParams()
    lib().cx_params_domain(C.byref(p), float(dr), f3(bbmin), f3(bbmax))
    return p


def params_fill_eps(p):
    lib().cx_params_fill_eps(C.byref(p))


def params_container(p, use, cmin=(0, 0, 0), cmax=(0, 0, 0)):
    lib().cx_params_container(C.byref(p), int(use), f3(cmin), f3(cmax))


def fill_dims(p):
    d = (I64 * 3)()
    lib().cx_fill_dims(C.byref(p), d)
    return np.array(list(d), dtype=np.int64)


def seed_index(p, spos):
    return int(lib().cx_seed_index(C.byref(p), f3(spos)))


def dist_unique_id():
    """128 bytes identifying a new NCCL communicator (ncclGetUniqueId); rank 0 calls it and ships the bytes to the other ranks"""
    buf = (C.c_char * 128)()
    rc = lib().cx_dist_unique_id(C.cast(buf, VP))
    if rc != 0:
        raise CxError(rc)
    return bytes(buf.raw)


def dist_init_local(contexts):
    """the given contexts (same process; one host thread per context must then drive it) form a group without NCCL"""
    arr = (VP * len(contexts))(*[c.h for c in contexts])
    rc = lib().cx_dist_init_local(arr, len(contexts))
    if rc != 0:
        raise CxError(rc)


def weld_host(corners, dr):
    corners = np.ascontiguousarray(corners, dtype=np.float32).reshape(-1, 3)
    out_v = np.zeros_like(corners)
    out_id = np.zeros(corners.shape[0], dtype=np.int32)
    nv = lib().cx_weld_host(corners.ctypes.data_as(C.POINTER(C.c_float)), corners.shape[0], float(dr),
                            out_v.ctypes.data_as(C.POINTER(C.c_float)), out_id.ctypes.data_as(C.POINTER(C.c_int32)))
    return out_v[:nv].copy(), out_id.reshape(-1, 3)


# ---- device context
class Context:
    """One cx_ctx per process / GPU.  Stage methods take torch CUDA tensors in the reference's layouts."""

    def __init__(self, device=0, use_torch_stream=True):
        import torch
        if not torch.cuda.is_available():
            raise CxError(-105, "no CUDA device visible (crixus_b200 has no CPU path)")
        self.L = lib()
        self.h = VP()
        rc = self.L.cx_create(int(device), C.byref(self.h))
        if rc != 0:
            raise CxError(rc)
        self.device = torch.device("cuda", int(device))
        if use_torch_stream:
            torch.cuda.set_device(self.device)
            self.L.cx_set_stream(self.h, VP(torch.cuda.current_stream(self.device).cuda_stream))

    def close(self):
        if self.h:
            self.L.cx_destroy(self.h)
            self.h = VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise CxError(rc, (self.L.cx_last_error(self.h) or b"").decode())

    @staticmethod
    def _p(t):
        return VP(t.data_ptr()) if t is not None else VP()

    def launches(self):
        return int(self.L.cx_kernel_launches(self.h))

    def measure_fp32_peak(self, reps=5):
        t = C.c_double(0)
        self._ck(self.L.cx_measure_fp32_peak(self.h, reps, C.byref(t)))
        return float(t.value)

    def synchronize(self):
        self._ck(self.L.cx_synchronize(self.h))

    def swap_normals(self, norm):
        self._ck(self.L.cx_swap_normals(self.h, self._p(norm), norm.shape[0]))

    def set_bound_elem(self, pos, norm, surf, ep, nbe, nvert, want_zstat=True):
        z = (C.c_float * 4)()
        self._ck(self.L.cx_set_bound_elem(self.h, self._p(pos), self._p(norm), self._p(surf), self._p(ep), nbe, nvert,
                                          z if want_zstat else None))
        return np.array(list(z), dtype=np.float32)

    def bin_segments(self, p, pre_pos, pos, pre_norm, norm, pre_ep, ep, pre_surf, surf, cell_idx, nbe, nvert):
        self._ck(self.L.cx_bin_segments(self.h, C.byref(p), self._p(pre_pos), self._p(pos), self._p(pre_norm), self._p(norm),
                                        self._p(pre_ep), self._p(ep), self._p(pre_surf), self._p(surf), self._p(cell_idx), nbe, nvert))

    def find_links(self, p, pos, nvert, newlink, idim):
        self._ck(self.L.cx_find_links(self.h, C.byref(p), self._p(pos), nvert, self._p(newlink), idim))

    def periodicity_links(self, ep, nbe, newlink):
        self._ck(self.L.cx_periodicity_links(self.h, self._p(ep), nbe, self._p(newlink)))

    def calc_trisize(self, ep, trisize, nbe):
        self._ck(self.L.cx_calc_trisize(self.h, self._p(ep), self._p(trisize), nbe))

    def identify_special_boundary(self, p, pos, ep, cell_idx, nvert, nbe, sbpos, sbep, sbnbe, sbid, sbi):
        self._ck(self.L.cx_identify_special_boundary(self.h, C.byref(p), self._p(pos), self._p(ep), self._p(cell_idx), nvert, nbe,
                                                     self._p(sbpos), self._p(sbep), sbnbe, self._p(sbid), int(sbi)))

    def calc_vert_volume(self, p, pos, norm, ep, vol, trisize, cell_idx, nvert, nbe, zero_open=False, v_begin=0, v_end=None):
        self._ck(self.L.cx_calc_vert_volume(self.h, C.byref(p), self._p(pos), self._p(norm), self._p(ep), self._p(vol),
                                            self._p(trisize), self._p(cell_idx), nvert, nbe, int(zero_open), v_begin,
                                            nvert if v_end is None else v_end))

    def calc_vert_volume_part(self, p, pos, norm, ep, vol, trisize, cell_idx, nvert, nbe, zero_open, chunk, nparts, part):
        self._ck(self.L.cx_calc_vert_volume_part(self.h, C.byref(p), self._p(pos), self._p(norm), self._p(ep), self._p(vol),
                                                 self._p(trisize), self._p(cell_idx), nvert, nbe, int(zero_open), int(chunk), int(nparts), int(part)))

    def fill_box(self, p, fpos, lo, hi):
        n = I64(0)
        self._ck(self.L.cx_fill_box(self.h, C.byref(p), self._p(fpos), f3(lo), f3(hi), C.byref(n)))
        return int(n.value)

    def fill_geometry(self, p, fpos, norm, ep, pos, nbe, cnbe, cell_idx, seed):
        n, r = I64(0), C.c_int32(0)
        self._ck(self.L.cx_fill_geometry(self.h, C.byref(p), self._p(fpos), self._p(norm), self._p(ep), self._p(pos), nbe, cnbe,
                                         self._p(cell_idx), seed, C.byref(n), C.byref(r)))
        self.last_rounds = int(r.value)
        return int(n.value)

    def fill_geometry_slab(self, p, fpos, norm, ep, pos, nbe, cnbe, cell_idx, seed, k_begin, k_end, first_call):
        n, r = I64(0), C.c_int32(0)
        self._ck(self.L.cx_fill_geometry_slab(self.h, C.byref(p), self._p(fpos), self._p(norm), self._p(ep), self._p(pos), nbe, cnbe,
                                              self._p(cell_idx), seed, k_begin, k_end, int(first_call), C.byref(n), C.byref(r)))
        self.last_rounds = int(r.value)
        return int(n.value)

    def fill_set_max_rounds(self, n):
        self._ck(self.L.cx_fill_set_max_rounds(self.h, int(n)))

    def fill_unfinished(self):
        return int(self.L.cx_fill_unfinished(self.h))

    def fill_set_resolve_mode(self, mode):
        """0 = record-driven directed tests at fine dr (default), 1 = the cell-driven kernel always"""
        self._ck(self.L.cx_fill_set_resolve_mode(self.h, int(mode)))

    def fill_resolve_slab(self, p, fpos, norm, ep, pos, nbe, cnbe, cell_idx, k_begin, k_end):
        """-> torch int32 view (zero-copy) of the 6 blocked bit planes owned by the context"""
        import torch
        ptr, n = VP(), I64(0)
        self._ck(self.L.cx_fill_resolve_slab(self.h, C.byref(p), self._p(fpos), self._p(norm), self._p(ep), self._p(pos), nbe, cnbe,
                                             self._p(cell_idx), k_begin, k_end, C.byref(ptr), C.byref(n)))

        class _Raw:
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": "<i4", "data": (int(ptr.value), False), "version": 2}
        return torch.as_tensor(_Raw(), device=self.device)

    def fill_resolve_part(self, p, fpos, norm, ep, pos, nbe, cnbe, cell_idx, nparts, part):
        import torch
        ptr, n = VP(), I64(0)
        self._ck(self.L.cx_fill_resolve_part(self.h, C.byref(p), self._p(fpos), self._p(norm), self._p(ep), self._p(pos), nbe, cnbe,
                                             self._p(cell_idx), int(nparts), int(part), C.byref(ptr), C.byref(n)))

        class _Raw:
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": "<i4", "data": (int(ptr.value), False), "version": 2}
        return torch.as_tensor(_Raw(), device=self.device)

    def fill_propagate(self, p, fpos, norm, ep, pos, nbe, cnbe, cell_idx, seed):
        n, r = I64(0), C.c_int32(0)
        self._ck(self.L.cx_fill_propagate(self.h, C.byref(p), self._p(fpos), self._p(norm), self._p(ep), self._p(pos), nbe, cnbe,
                                          self._p(cell_idx), seed, C.byref(n), C.byref(r)))
        self.last_rounds = int(r.value)
        return int(n.value)

    def check_triangle_size(self, corners, dr):
        """corners: torch float32 (3*ntri, 3) on the device -> (failures, min distance, max distance)"""
        n, lo, hi = I64(0), C.c_float(0), C.c_float(0)
        self._ck(self.L.cx_check_triangle_size(self.h, self._p(corners), corners.shape[0] // 3, float(dr), C.byref(n), C.byref(lo), C.byref(hi)))
        return int(n.value), float(lo.value), float(hi.value)

    def weld_device(self, corners, dr):
        """corners: torch float32 (ncorner, 3) on the device -> (pos float4 (nvert,4), ep int4 (nbe,4), bbmin, bbmax)"""
        import torch
        n = corners.shape[0]
        lo, hi = F3c(), F3c()
        self._ck(self.L.cx_bbox_device(self.h, self._p(corners), n, lo, hi))
        pos4 = torch.zeros((n, 4), dtype=torch.float32, device=corners.device)
        ep4 = torch.zeros((n // 3, 4), dtype=torch.int32, device=corners.device)
        nv = I64(0)
        self._ck(self.L.cx_weld_device(self.h, self._p(corners), n, float(dr), lo, hi, self._p(pos4), self._p(ep4), C.byref(nv)))
        return pos4[: int(nv.value)], ep4, np.array(list(lo), dtype=np.float32), np.array(list(hi), dtype=np.float32)

    def preprocess_host(self, job):
        self._ck(self.L.cx_preprocess_host(self.h, C.byref(job)))

    # ---- multi-GPU (include/crixus_b200.h "multi-GPU"): the communicator lives inside the library
    def dist_init(self, id128, world, rank):
        """joins the NCCL communicator identified by the 128 bytes of cx_dist_unique_id (created on rank 0, shipped by the caller)"""
        buf = (C.c_char * 128).from_buffer_copy(bytes(id128))
        self._ck(self.L.cx_dist_init(self.h, C.cast(buf, VP), int(world), int(rank)))

    def dist_finalize(self):
        self.L.cx_dist_finalize(self.h)

    def dist_set_fill_mode(self, mode):
        self._ck(self.L.cx_dist_set_fill_mode(self.h, int(mode)))

    def dist_world(self):
        return int(self.L.cx_dist_world(self.h))

    def dist_rank(self):
        return int(self.L.cx_dist_rank(self.h))

    def preprocess_dist(self, job):
        self._ck(self.L.cx_preprocess_dist(self.h, C.byref(job)))

    def dist_calc_vert_volume(self, p, pos, norm, ep, vol, trisize, cell_idx, nvert, nbe, zero_open=False):
        self._ck(self.L.cx_dist_calc_vert_volume(self.h, C.byref(p), self._p(pos), self._p(norm), self._p(ep), self._p(vol),
                                                 self._p(trisize), self._p(cell_idx), nvert, nbe, int(zero_open)))

    def dist_fill_geometry(self, p, fpos, norm, ep, pos, nbe, cnbe, cell_idx, seed):
        n, r = I64(0), C.c_int32(0)
        self._ck(self.L.cx_dist_fill_geometry(self.h, C.byref(p), self._p(fpos), self._p(norm), self._p(ep), self._p(pos), nbe, cnbe,
                                              self._p(cell_idx), seed, C.byref(n), C.byref(r)))
        self.last_rounds = int(r.value)
        return int(n.value)

    def hash_u32(self, t):
        """(order-independent 64-bit checksum, popcount) of a 4-byte-element CUDA tensor, computed on the device"""
        out = (C.c_uint64 * 2)()
        self._ck(self.L.cx_hash_u32(self.h, self._p(t), t.numel() * t.element_size() // 4, out))
        return int(out[0]), int(out[1])

    def assemble_records_device(self, job, nfluid_for_indices=None):
        """cx_assemble_records_device on the device arrays a cx_preprocess_host call left behind -> torch uint8 (n, 64)"""
        import torch
        cap = int(job.nfluid) + int(job.nvert) + int(job.nbe)
        rec = torch.empty((max(cap, 1), 64), dtype=torch.uint8, device=self.device)
        n = I64(0)
        self._ck(self.L.cx_assemble_records_device(self.h, C.byref(job.params_fill), job.dimg, VP(job.d_fpos), VP(job.d_pos), VP(job.d_norm),
                                                   VP(job.d_ep), VP(job.d_surf), VP(job.d_vol), VP(job.d_sbid), int(job.nvert), int(job.nbe),
                                                   int(job.nfluid if nfluid_for_indices is None else nfluid_for_indices), VP(rec.data_ptr()),
                                                   cap, C.byref(n)))
        self.synchronize()
        return rec[: int(n.value)]

    def stream_h5sph_device(self, job, filename, chunk_items=0, nfluid_for_indices=None):
        """cx_stream_h5sph_device on the device arrays a cx_preprocess_host call left behind -> records written"""
        n = I64(0)
        self._ck(self.L.cx_stream_h5sph_device(self.h, C.byref(job.params_fill), job.dimg, VP(job.d_fpos), VP(job.d_pos), VP(job.d_norm),
                                               VP(job.d_ep), VP(job.d_surf), VP(job.d_vol), int(job.nvert), int(job.nbe),
                                               int(job.nfluid if nfluid_for_indices is None else nfluid_for_indices),
                                               str(filename).encode(), int(chunk_items), C.byref(n)))
        return int(n.value)

    def write_records_device(self, rec, filename, fmt="vtu"):
        fn = self.L.cx_write_h5sph_records_device if fmt == "h5sph" else self.L.cx_write_vtu_records_device
        self._ck(fn(self.h, VP(rec.data_ptr()), int(rec.shape[0]), str(filename).encode()))
