"""Test-side adapter: exposes the CUDA path (C-ABI through crixus_b200.api) with the same stage interface as
oracle.orc.Oracle, so oracle.orc.run_pipeline can drive either.  numpy in / numpy out; torch only carries the bytes."""
import numpy as np
import torch

from crixus_b200 import api


class CudaBackend:
    name = "cuda"

    def __init__(self, device=0):
        self.ctx = api.Context(device)
        self.dev = self.ctx.device
        self.last_sweeps = 0

    def _d(self, a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    # ---- params: the product's own host glue
    def params_domain(self, dr, bbmin, bbmax):
        return api.params_domain(dr, bbmin, bbmax)

    def params_fill_eps(self, p):
        api.params_fill_eps(p)

    def params_container(self, p, use, cmin=(0, 0, 0), cmax=(0, 0, 0)):
        api.params_container(p, use, cmin, cmax)

    def fill_dims(self, p):
        return api.fill_dims(p)

    def seed_index(self, p, spos):
        return api.seed_index(p, spos)

    def weld(self, corners, dr):
        return api.weld_host(corners, dr)

    # ---- stages
    def swap_normals(self, p, norm):
        d = self._d(norm)
        self.ctx.swap_normals(d)
        norm[...] = d.cpu().numpy()

    def set_bound_elem(self, p, pos, norm, surf, ep, nbe, nvert):
        dpos, dnorm, dsurf, dep = self._d(pos), self._d(norm), self._d(surf), self._d(ep)
        z = self.ctx.set_bound_elem(dpos, dnorm, dsurf, dep, nbe, nvert)
        pos[...] = dpos.cpu().numpy()
        surf[...] = dsurf.cpu().numpy()
        ep[...] = dep.cpu().numpy()
        return z

    def bin_segments(self, p, pre_pos, pre_norm, pre_ep, pre_surf, nbe, nvert):
        d = [self._d(x) for x in (pre_pos, pre_norm, pre_ep, pre_surf)]
        pos, norm, ep, surf = (torch.zeros_like(x) for x in d)
        cell_idx = torch.zeros(p.gridn[3] + 1, dtype=torch.int32, device=self.dev)
        self.ctx.bin_segments(p, d[0], pos, d[1], norm, d[2], ep, d[3], surf, cell_idx, nbe, nvert)
        return tuple(x.cpu().numpy() for x in (pos, norm, ep, surf, cell_idx))

    def find_links(self, p, pos, nvert, newlink, idim):
        dpos, dl = self._d(pos), self._d(newlink)
        try:
            self.ctx.find_links(p, dpos, nvert, dl, idim)
            missing = 0
        except api.CxError as e:
            if e.code != -9:
                raise
            missing = 1
        pos[...] = dpos.cpu().numpy()
        newlink[...] = dl.cpu().numpy()
        return missing

    def periodicity_links(self, p, pos, ep, nvert, nbe, newlink, idim):
        dep, dl = self._d(ep), self._d(newlink)
        self.ctx.periodicity_links(dep, nbe, dl)
        ep[...] = dep.cpu().numpy()

    def calc_trisize(self, p, ep, nvert, nbe):
        t = torch.zeros(nvert, dtype=torch.int32, device=self.dev)
        self.ctx.calc_trisize(self._d(ep), t, nbe)
        return t.cpu().numpy()

    def calc_vert_volume(self, p, pos, norm, ep, trisize, cell_idx, nvert, nbe, zero_open=False, v0=0, v1=None):
        vol = torch.zeros(nvert, dtype=torch.float32, device=self.dev)
        self.ctx.calc_vert_volume(p, self._d(pos), self._d(norm), self._d(ep), vol, self._d(trisize), self._d(cell_idx), nvert, nbe,
                                  zero_open, v0, v1)
        return vol.cpu().numpy()

    def identify_special_boundary(self, p, pos, ep, nvert, nbe, sbpos, sbep, sbid, sbi, cell_idx=None):
        d = self._d(sbid)
        self.ctx.identify_special_boundary(p, self._d(pos), self._d(ep), self._d(cell_idx), nvert, nbe, self._d(sbpos), self._d(sbep),
                                           sbep.shape[0], d, sbi)
        sbid[...] = d.cpu().numpy()

    def fill_box(self, p, fpos, lo, hi):
        d = self._d(fpos.view(np.int32))
        n = self.ctx.fill_box(p, d, lo, hi)
        fpos[...] = d.cpu().numpy().view(np.uint32)
        return n

    def fill_geometry(self, p, fpos, norm, ep, pos, nbe, cnbe, cell_idx, seed):
        d = self._d(fpos.view(np.int32))
        n = self.ctx.fill_geometry(p, d, self._d(norm), self._d(ep), self._d(pos), nbe, cnbe, self._d(cell_idx), seed)
        self.last_sweeps = self.ctx.last_rounds
        fpos[...] = d.cpu().numpy().view(np.uint32)
        return n
