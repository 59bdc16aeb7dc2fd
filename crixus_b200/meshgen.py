"""Deterministic synthetic meshes + binary-STL I/O for the workloads named in BASELINE.json.

All generators build *indexed* meshes (shared vertices are bit-identical float32), orient the facet normals
OUT of the fluid (the convention of the shipped SPHERIC-2 case, which then needs ``swap_normals=true``,
/root/reference/README.md:136) and satisfy the reference's mesh precondition: every vertex-barycentre
distance <= dr*(1-1e-4) (/root/reference/resources/test-triangle-size.c:24,47-56).

Binary STL layout follows the reference's reader, /root/reference/src/crixus.cu:126-213:
80-byte header (must not start with "solid"), uint32 count, count x {12 float32, uint16}.
"""
from __future__ import annotations

import numpy as np

RNG_SEED = 20140327


# ------------------------------------------------------------------------------------------------ STL I/O
def write_stl(path, tris, normals):
    """tris: (n,3,3) float32 corner coordinates; normals: (n,3) float32 facet normals (stored verbatim)."""
    tris = np.ascontiguousarray(tris, dtype="<f4")
    normals = np.ascontiguousarray(normals, dtype="<f4")
    n = tris.shape[0]
    rec = np.zeros(n, dtype=[("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")])
    rec["n"] = normals
    rec["v"] = tris
    with open(path, "wb") as f:
        f.write(b"crixus_b200 synthetic mesh".ljust(80, b" "))
        f.write(np.uint32(n).tobytes())
        f.write(rec.tobytes())


def read_stl(path):
    with open(path, "rb") as f:
        head = f.read(80)
        if head[:5] == b"solid":
            raise ValueError("ASCII STL (STL_NOT_BINARY)")
        n = int(np.frombuffer(f.read(4), dtype="<u4")[0])
        rec = np.frombuffer(f.read(50 * n), dtype=[("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")])
    if rec.shape[0] != n:
        raise ValueError("READ_ERROR: facet count mismatch")
    return rec["v"].copy(), rec["n"].copy()


# ------------------------------------------------------------------------------------------------ helpers
def _axis(breaks, h):
    """1-D coordinate array through all breakpoints, every interval split into equal parts <= h (float32)."""
    out = [np.float64(breaks[0])]
    for a, b in zip(breaks[:-1], breaks[1:]):
        n = max(int(np.ceil((b - a) / h - 1e-9)), 1)
        out.extend(a + (b - a) * (np.arange(1, n + 1) / n))
    return np.asarray(out, dtype=np.float64).astype(np.float32)


class _Builder:
    """Accumulates axis-aligned faces as right-triangle grids (vectorised); shared vertices are merged at the end by
    their float32 bit pattern, so faces built from the same coordinate arrays are conforming and watertight."""

    def __init__(self):
        self.verts = []
        self.tris = []
        self.normals = []
        self.nv = 0

    def quad_face(self, axis, const, U, V, normal_sign, keep=None):
        """Axis-aligned face at coordinate ``const`` on ``axis``; U, V are the coordinate arrays of the two other
        axes in cyclic order (axis+1, axis+2).  Each quad gets two right triangles wound to match the normal
        ``normal_sign`` * e_axis, emitted quad by quad with iu outer and iv inner.  keep(iu, iv) -> bool array drops
        quads (holes)."""
        a1, a2 = (axis + 1) % 3, (axis + 2) % 3
        U = np.asarray(U, dtype=np.float32)
        V = np.asarray(V, dtype=np.float32)
        nu, nv = len(U), len(V)
        if nu < 2 or nv < 2:
            return
        P = np.zeros((nu, nv, 3), dtype=np.float32)
        P[..., axis] = np.float32(const)
        P[..., a1] = U[:, None]
        P[..., a2] = V[None, :]
        idx = (np.arange(nu)[:, None] * nv + np.arange(nv)[None, :]) + self.nv
        a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
        if normal_sign > 0:   # (a,b,c),(a,c,d) is counter-clockwise seen from +axis
            t = np.stack([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], 2)
        else:
            t = np.stack([np.stack([a, c, b], -1), np.stack([a, d, c], -1)], 2)
        if keep is not None:
            iu, iv = np.meshgrid(np.arange(nu - 1), np.arange(nv - 1), indexing="ij")
            t = t[np.asarray(keep(iu, iv), dtype=bool)]
        t = t.reshape(-1, 3)
        nrm = np.zeros(3, dtype=np.float32)
        nrm[axis] = normal_sign
        self.verts.append(P.reshape(-1, 3))
        self.tris.append(t)
        self.normals.append(np.tile(nrm, (t.shape[0], 1)))
        self.nv += nu * nv

    def finish(self):
        v = np.concatenate(self.verts)
        t = np.concatenate(self.tris)
        n = np.concatenate(self.normals)
        key = np.ascontiguousarray(v).view(np.dtype((np.void, 12))).ravel()
        _, first, inv = np.unique(key, return_index=True, return_inverse=True)
        return v[first].astype(np.float32), inv[t].astype(np.int32), n.astype(np.float32)


def soup(verts, tris):
    """(nt,3,3) corner coordinates of an indexed mesh."""
    return verts[tris]


def geometric_normals(verts, tris, flip=False):
    v = verts.astype(np.float64)
    n = np.cross(v[tris[:, 1]] - v[tris[:, 0]], v[tris[:, 2]] - v[tris[:, 0]])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return (-n if flip else n).astype(np.float32)


def triangle_size_ok(verts, tris, dr):
    """The criterion of resources/test-triangle-size.c: max vertex-barycentre distance <= dr - 1e-4*dr."""
    c = verts[tris].astype(np.float64)
    bary = c.mean(axis=1, keepdims=True)
    return float(np.sqrt(((c - bary) ** 2).sum(-1)).max()) <= dr - 1e-4 * dr


def jitter_interior(verts, tris, normals, amp, rng=None):
    """In-plane jitter of vertices that touch facets of ONE normal direction only (face-interior vertices)."""
    rng = rng or np.random.default_rng(RNG_SEED)
    nv = verts.shape[0]
    first = np.full((nv, 3), np.nan, dtype=np.float32)
    same = np.ones(nv, dtype=bool)
    for k in range(3):
        for t in range(tris.shape[0]):
            v = tris[t, k]
            if np.isnan(first[v, 0]):
                first[v] = normals[t]
            elif not np.array_equal(first[v], normals[t]):
                same[v] = False
    out = verts.copy()
    d = rng.uniform(-amp, amp, size=(nv, 3)).astype(np.float32)
    d -= (d * first).sum(1, keepdims=True) * first  # remove the normal component
    out[same] += d[same]
    return out


def subdivide(verts, tris, normals, levels=1):
    """Midpoint (1->4) subdivision; facet normals are inherited, midpoints are float32((a+b)/2) per edge."""
    for _ in range(levels):
        e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])
        es = np.sort(e, axis=1)
        uniq, inv = np.unique(es, axis=0, return_inverse=True)
        mid = ((verts[uniq[:, 0]].astype(np.float64) + verts[uniq[:, 1]].astype(np.float64)) * 0.5).astype(np.float32)
        nv = verts.shape[0]
        nt = tris.shape[0]
        m01, m12, m20 = nv + inv[:nt], nv + inv[nt:2 * nt], nv + inv[2 * nt:]
        a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
        tris = np.concatenate([np.stack([a, m01, m20], 1), np.stack([m01, b, m12], 1), np.stack([m20, m12, c], 1),
                               np.stack([m01, m12, m20], 1)]).astype(np.int32)
        normals = np.concatenate([normals] * 4)
        verts = np.concatenate([verts, mid])
    return verts, tris, normals


# ------------------------------------------------------------------------------------------------ geometries
def plate(n=10, h=0.1):
    """KAT-plate of SURVEY.md App. D.3: (n+1)^2 vertices (i*h, j*h, 0), two triangles (a,b,d),(a,d,c) per quad,
    STL normals (0,0,1).  Vertex id j*(n+1)+i."""
    hf = np.float32(h)
    verts = np.array([[np.float32(i) * hf, np.float32(j) * hf, 0.0] for j in range(n + 1) for i in range(n + 1)], dtype=np.float32)
    tris = []
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i
            b, c = a + 1, a + n + 1
            d = c + 1
            tris += [(a, b, d), (a, d, c)]
    tris = np.asarray(tris, dtype=np.int32)
    normals = np.tile(np.array([0, 0, 1], dtype=np.float32), (tris.shape[0], 1))
    return verts, tris, normals


def box_tank(lo, hi, h, open_top=False, inward=False, obstacle=None):
    """Axis-aligned tank [lo,hi] meshed with right-triangle grids of spacing <= h.  Normals point out of the
    fluid (outward) unless inward=True.  obstacle=(olo, ohi): a box standing on the floor (its footprint is cut
    out of the floor; its sides/top face the fluid)."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    br = [[lo[a], hi[a]] for a in range(3)]
    if obstacle is not None:
        olo, ohi = np.asarray(obstacle[0], dtype=np.float64), np.asarray(obstacle[1], dtype=np.float64)
        for a in range(3):
            br[a] = sorted(set(br[a] + [olo[a], ohi[a]]))
    ax = [_axis(br[a], h) for a in range(3)]
    s = -1.0 if inward else 1.0
    b = _Builder()

    def idx(a, x):
        return int(np.argmin(np.abs(ax[a].astype(np.float64) - x)))

    hole = None
    if obstacle is not None:
        ix0, ix1, iy0, iy1 = idx(0, olo[0]), idx(0, ohi[0]), idx(1, olo[1]), idx(1, ohi[1])
        hole = lambda iu, iv: ~((ix0 <= iu) & (iu < ix1) & (iy0 <= iv) & (iv < iy1))  # noqa: E731
    # z faces: axis 2 -> (U,V) = (x,y)
    b.quad_face(2, ax[2][0], ax[0], ax[1], -s, keep=hole)
    if not open_top:
        b.quad_face(2, ax[2][-1], ax[0], ax[1], +s)
    # x faces: axis 0 -> (U,V) = (y,z);  y faces: axis 1 -> (U,V) = (z,x)
    b.quad_face(0, ax[0][0], ax[1], ax[2], -s)
    b.quad_face(0, ax[0][-1], ax[1], ax[2], +s)
    b.quad_face(1, ax[1][0], ax[2], ax[0], -s)
    b.quad_face(1, ax[1][-1], ax[2], ax[0], +s)
    if obstacle is not None:
        iz1 = idx(2, ohi[2])
        X, Y, Z = ax[0][ix0:ix1 + 1], ax[1][iy0:iy1 + 1], ax[2][0:iz1 + 1]
        # obstacle surfaces: fluid is outside the box, "out of the fluid" = into the box
        b.quad_face(2, ax[2][iz1], X, Y, -s)
        b.quad_face(0, X[0], Y, Z, +s)
        b.quad_face(0, X[-1], Y, Z, -s)
        b.quad_face(1, Y[0], Z, X, +s)
        b.quad_face(1, Y[-1], Z, X, -s)
    return b.finish()


def spheric2(dr=0.018333, hfac=1.25):
    """SPHERIC test 2 (dam break with obstacle): tank [0,3.22]x[0,1]x[0,1], open top, box obstacle
    0.161 x 0.403 x 0.161 on the floor (the shipped STL is missing from the reference tree)."""
    return box_tank((0, 0, 0), (3.22, 1.0, 1.0), hfac * dr, open_top=True,
                    obstacle=((2.3955, 0.2985, 0.0), (2.5565, 0.7015, 0.161)))


def spheric2_fshape():
    """The 4-triangle free-surface shape shipped as resources/spheric2_fshape.stl (lid z=0.551 over x in [0,1.228],
    wall x=1.228 for z in [0,0.551]), regenerated (decoded in SURVEY.md §4)."""
    x1, z1 = np.float32(1.228), np.float32(0.551)
    v = np.array([[0, 0, z1], [x1, 0, z1], [x1, 1, z1], [0, 1, z1], [x1, 0, 0], [x1, 1, 0]], dtype=np.float32)
    tris = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 2], [4, 2, 1]], dtype=np.int32)
    return v, tris, geometric_normals(v, tris)


def channel(n=24, dr=0.05, hfac=1.25, height=None):
    """Periodic channel (x,y periodic): floor z=0 and lid z=H over [0,L]^2, no side walls; vertices on the max
    faces are exact mirrors of the min faces."""
    h = hfac * dr
    L = n * h
    H = height if height is not None else max(6, n // 4) * h
    X = _axis([0.0, L], h)
    b = _Builder()
    b.quad_face(2, np.float32(0.0), X, X, -1.0)
    b.quad_face(2, np.float32(H), X, X, +1.0)
    return b.finish()


def icosphere(center, radius, levels):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                  [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(levels):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        uniq, inv = np.unique(np.sort(e, axis=1), axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        nv, nt = v.shape[0], f.shape[0]
        m01, m12, m20 = nv + inv[:nt], nv + inv[nt:2 * nt], nv + inv[2 * nt:]
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        f = np.concatenate([np.stack([a, m01, m20], 1), np.stack([m01, b, m12], 1), np.stack([m20, m12, c], 1),
                            np.stack([m01, m12, m20], 1)]).astype(np.int32)
        v = np.concatenate([v, mid])
    verts = (np.asarray(center, dtype=np.float64) + radius * v).astype(np.float32)
    return verts, f


def sphere_in_tank(side_cells=40, dr=0.025, hfac=1.0, sphere_levels=3):
    """Closed cubic tank of side side_cells*dr with a sphere of radius side/4 at the centre.
    Fluid fills the space between sphere and wall; normals point out of the fluid."""
    side = side_cells * dr
    tv, tt, tn = box_tank((0, 0, 0), (side, side, side), hfac * dr)
    sv, st = icosphere((side / 2,) * 3, side / 4, sphere_levels)
    sn = geometric_normals(sv, st, flip=True)  # out of the fluid = into the sphere
    verts = np.concatenate([tv, sv])
    tris = np.concatenate([tt, st + tv.shape[0]]).astype(np.int32)
    return verts, tris, np.concatenate([tn, sn])
