"""The parity cases shared by the oracle tests, the GPU tests and tests/golden/make_golden.py.  Each case is a small
version of one BASELINE.json config, sized so that the CPU oracle finishes in seconds."""
import numpy as np

from crixus_b200 import meshgen as mg


def _lcg_jitter(verts, tris, normals, amp, amp_n=0.0):
    rng = np.random.default_rng(mg.RNG_SEED)
    v = mg.jitter_interior(verts, tris, normals, amp, rng)
    if amp_n:
        # out-of-plane jitter of interior vertices (the D.4 sensitivity mesh of SURVEY.md): z only, plate case
        inner = ~np.isclose(v, verts).all(1)
        v[inner, 2] += rng.uniform(-amp_n, amp_n, size=int(inner.sum())).astype(np.float32)
    return v


# the reference program refuses to run without a fill_0 section (crixus.cu:737,759-763): give the open plates a
# container above the plate and a small box fill (which also exercises fill_fluid)
_PLATE_FILL = dict(container=((0, 0, 0), (1.0, 1.0, 0.2)), fills=[("box", (0.4, 0.4, 0.08), (0.6, 0.6, 0.12))])


def case(name):
    """-> dict(tris, normals, dr, swap_normals, periodic, zero_open, container, fills, fshape)"""
    c = dict(swap_normals=False, periodic=(False, False, False), zero_open=False, container=None, fills=[], fshape=None, special=[])
    if name == "plate":            # KAT-plate, SURVEY.md App. D.3
        v, t, n = mg.plate(10, 0.1)
        c.update(dr=0.08, **_PLATE_FILL)
    elif name == "plate_jitter":   # D.4: lifted, jittered plate -> generic (non axis-aligned) plane tests
        v, t, n = mg.plate(24, 0.1)
        v = v.copy()
        v[:, 2] += np.float32(0.013)
        v = _lcg_jitter(v, t, n, 0.01, 0.0025)
        n = mg.geometric_normals(v, t)
        c.update(dr=0.08, **_PLATE_FILL)
    elif name == "plate_zero_open":
        v, t, n = mg.plate(6, 0.1)
        c.update(dr=0.08, zero_open=True, container=((0, 0, 0), (0.6, 0.6, 0.2)), fills=[("box", (0.2, 0.2, 0.08), (0.4, 0.4, 0.12))])
    elif name == "fan20":          # one vertex with 20 triangles (> the 16-triangle fast path, < trimax = 50)
        k = 20
        ang = 2 * np.pi * np.arange(k) / k
        v = np.concatenate([[[0.5, 0.5, 0.0]], np.stack([0.5 + 0.07 * np.cos(ang), 0.5 + 0.07 * np.sin(ang), np.zeros(k)], 1)]).astype(np.float32)
        t = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)], dtype=np.int32)
        n = np.tile(np.array([0, 0, 1], dtype=np.float32), (k, 1))
        c.update(dr=0.08, **_PLATE_FILL)
    elif name == "cube":           # KAT-cube, SURVEY.md App. D.3
        v, t, n = mg.box_tank((0, 0, 0), (1, 1, 1), 0.1, inward=True)
        c.update(dr=0.08, fills=[("geometry", (0.5, 0.5, 0.5), 0.08)])
    elif name == "cube_jitter":    # D.5
        v, t, n = mg.box_tank((0, 0, 0), (1, 1, 1), 0.1, inward=True)
        rng = np.random.default_rng(mg.RNG_SEED)
        interior = ((v > 1e-6) & (v < 1 - 1e-6)).sum(1) == 2   # face-interior vertices
        v = v.copy()
        ids = np.nonzero(interior)[0]
        ids = ids[np.lexsort((v[ids, 2], v[ids, 1], v[ids, 0]))]   # independent of the generator's vertex numbering
        v[ids] += rng.uniform(-0.003, 0.003, size=(ids.size, 3)).astype(np.float32)
        n = mg.geometric_normals(v, t)
        ctr = v[t].mean(1) - 0.5
        flip = (n * ctr).sum(1) > 0   # make normals point inward
        n[flip] *= -1
        c.update(dr=0.08, fills=[("geometry", (0.5, 0.5, 0.5), 0.08)])
    elif name == "cube_h05":       # h = dr/2: voronoi planes on voxel planes, edge planes through the faces of the sampling cube
        v, t, n = mg.box_tank((0, 0, 0), (1, 1, 1), 0.05, inward=True)
        c.update(dr=0.1, fills=[("geometry", (0.5, 0.5, 0.5), 0.1)])
    elif name == "cube_tiny":      # dr = 2e-4: eps/(|c| dr/40) > 1, several voxels of a line lie within eps of a voronoi plane
        dr = 2e-4
        v, t, n = mg.box_tank((0, 0, 0), (20 * dr, 20 * dr, 20 * dr), 0.625 * dr, inward=True)
        rng = np.random.default_rng(mg.RNG_SEED)
        interior = ((v > 1e-9) & (v < 20 * dr - 1e-9)).sum(1) == 2
        v = v.copy()
        ids = np.nonzero(interior)[0]
        ids = ids[np.lexsort((v[ids, 2], v[ids, 1], v[ids, 0]))]
        v[ids] += rng.uniform(-0.03 * dr, 0.03 * dr, size=(ids.size, 3)).astype(np.float32)
        n = mg.geometric_normals(v, t)
        ctr = v[t].mean(1) - 10 * dr
        n[(n * ctr).sum(1) > 0] *= -1
        c.update(dr=dr, fills=[("geometry", (10 * dr, 10 * dr, 10 * dr), dr)])
    elif name == "spheric2_mini":  # config 1 at a coarser dr: swap_normals, container, box + geometry fill, fshape
        dr = 0.06
        v, t, n = mg.spheric2(dr=dr)
        fv, ft, fn = mg.spheric2_fshape()
        c.update(dr=dr, swap_normals=True, container=((0, 0, 0), (1.5, 1.0, 0.6)),
                 fills=[("box", (0.2, 0.2, 0.2), (0.4, 0.4, 0.4)), ("geometry", (0.5, 0.5, 0.5), dr)], fshape=(fv[ft], fn))
    elif name == "channel":        # config 3 in small: x/y periodic floor+lid
        dr = 0.05
        v, t, n = mg.channel(n=16, dr=dr, height=10 * 1.25 * dr)
        L = 16 * 1.25 * dr
        c.update(dr=dr, swap_normals=True, periodic=(True, True, False),
                 fills=[("geometry", (L / 2, L / 2, 5 * 1.25 * dr), dr)])
    elif name == "channel_yz":     # the channel with its axes permuted: periodic in y and z (wrap of the z-columns in the fill,
        dr = 0.05                  # periodic unwrap along z in calc_vert_volume, find_links for idim = 1, 2)
        v, t, n = mg.channel(n=16, dr=dr, height=10 * 1.25 * dr)
        v, n = np.ascontiguousarray(v[:, [2, 0, 1]]), np.ascontiguousarray(n[:, [2, 0, 1]])
        L = 16 * 1.25 * dr
        c.update(dr=dr, swap_normals=True, periodic=(False, True, True),
                 fills=[("geometry", (5 * 1.25 * dr, L / 2, L / 2), dr)])
    elif name == "cube_fs_bad":    # free-surface triangles whose stored normal disagrees with their plane: nothing can be culled,
        v, t, n = mg.box_tank((0, 0, 0), (1, 1, 1), 0.1, inward=True)      # every lattice cell takes the full test
        fs_t = np.array([[[0, 0, 0.5], [1, 0, 0.5], [1, 1, 0.5]], [[0, 0, 0.5], [1, 1, 0.5], [0, 1, 0.5]]], dtype=np.float32)
        fs_n = np.array([[0, 0.6, 0.8], [0, 0, 0.5]], dtype=np.float32)
        c.update(dr=0.08, fills=[("geometry", (0.5, 0.5, 0.25), 0.08)], fshape=(fs_t, fs_n))
    elif name == "sphere_tank":    # config 4 in small: curved, non axis-aligned triangles
        dr = 0.05
        v, t, n = mg.sphere_in_tank(side_cells=20, dr=dr, hfac=1.0, sphere_levels=2)
        c.update(dr=dr, swap_normals=True, fills=[("geometry", (0.1, 0.1, 0.1), dr)])
    elif name == "cube_sb":        # special-boundary grids (crixus.cu:469-617): inlet patch on x=0, slanted fan on the lid, and a
        v, t, n = mg.box_tank((0, 0, 0), (1, 1, 1), 0.1, inward=True)   # third grid overlapping the first (later ids overwrite)
        def rect_x0(y0, y1, z0, z1):
            return np.array([[[0, y0, z0], [0, y1, z0], [0, y1, z1]], [[0, y0, z0], [0, y1, z1], [0, y0, z1]]], dtype=np.float32)
        ring = np.array([[0.3, 0.3, 1], [0.8, 0.35, 1], [0.7, 0.8, 1], [0.25, 0.7, 1]], dtype=np.float32)
        ctr = np.array([0.5, 0.5, 1], dtype=np.float32)
        fan = np.array([[ctr, ring[i], ring[(i + 1) % 4]] for i in range(4)], dtype=np.float32)
        c.update(dr=0.08, fills=[("geometry", (0.5, 0.5, 0.5), 0.08)], special=[rect_x0(0.2, 0.6, 0.2, 0.5), fan, rect_x0(0.5, 0.8, 0.2, 0.5)])
    elif name == "sphere_sb":      # special-boundary grid coincident with curved mesh triangles (the upper half of the sphere):
        dr = 0.05                  # plane test |n.(b-v0)| <= eps*|nx+ny+nz| on un-normalised normals, marginal near nx+ny+nz = 0
        v, t, n = mg.sphere_in_tank(side_cells=20, dr=dr, hfac=1.0, sphere_levels=2)
        tr = v[t]
        mid = 0.5 * (v.min(0) + v.max(0))
        on_sphere = (np.abs(tr - mid).max(axis=(1, 2)) < 0.45 * float((v.max(0) - v.min(0)).max()))
        upper = on_sphere & (tr[:, :, 2].mean(1) > mid[2])
        c.update(dr=dr, swap_normals=True, fills=[("geometry", (0.1, 0.1, 0.1), dr)], special=[np.ascontiguousarray(tr[upper])])
    elif name == "plate_sb_stress":   # adversarial special-boundary grids (special_stress below), incl. the rounding-sensitive kinds
        v, t, n = mg.plate(24, 0.1)
        v = v.copy()
        v[:, 2] += np.float32(0.013)
        v = _lcg_jitter(v, t, n, 0.01, 0.0025)
        n = mg.geometric_normals(v, t)
        sb = special_stress(v[t], 7, n=240)
        c.update(dr=0.08, special=[sb[:120], sb[120:]], **_PLATE_FILL)
    elif name == "tank_mm":        # a mesh in MILLIMETRES: eps = 1e-5 x extent = 0.2 is absolute, but enters the segment-triangle
        dr = 500.0                 # test as a relative tolerance (r <= 1 + eps; t1, t2 >= -eps in barycentric units): the free
        v, t, n = mg.box_tank((0, 0, 0), (20000, 15000, 10000), dr, inward=True)   # surface blocks cells up to eps |u| = 7 dr
        fs_t = np.array([[[0, 0, 5000], [9750, 0, 5000], [9750, 15000, 5000]],      # beyond its edge; it spans half the tank
                         [[0, 0, 5000], [9750, 15000, 5000], [0, 15000, 5000]],
                         [[12000, 2000, 2000], [16000, 2000, 2600], [14000, 9000, 2300]]], dtype=np.float32)   # + a slanted one
        fs_n = mg.geometric_normals(fs_t.reshape(-1, 3), np.arange(9, dtype=np.int32).reshape(3, 3))
        c.update(dr=dr, fills=[("geometry", (2250.0, 2250.0, 2250.0), dr)], fshape=(fs_t, fs_n))
    else:
        raise KeyError(name)
    c.update(name=name, tris=np.ascontiguousarray(v[t], dtype=np.float32), normals=np.ascontiguousarray(n, dtype=np.float32),
             dr=np.float32(c["dr"]))
    return c


ALL = ["plate", "plate_jitter", "plate_zero_open", "fan20", "cube", "cube_jitter", "cube_h05", "cube_tiny", "spheric2_mini", "channel", "channel_yz", "cube_fs_bad", "sphere_tank", "cube_sb", "sphere_sb", "plate_sb_stress"]


# cases without a golden fixture of the reference CUDA binary (pinned against the reference's host build instead)
EXTRA = ["tank_mm"]


def run(B, c, stages=("volume", "fill")):
    from oracle import orc
    return orc.run_pipeline(B, c["tris"], c["normals"], c["dr"], swap_normals=c["swap_normals"], periodic=c["periodic"],
                            zero_open=c["zero_open"], container=c["container"], fills=c["fills"], fshape=c["fshape"], stages=stages,
                            special=c.get("special", ()))


def special_stress(tris, seed, n=160, skip=()):
    """Adversarial special-boundary grid for a mesh: triangles built around segment barycentres so
    that the barycentre sits at chosen barycentric coordinates (corner, edge, just inside / outside the eps bands of
    segInTri, crixus_d.cu:1065,1081), lifted out of the plane by multiples of eps, with edge lengths from 0.02 to 3 domain
    sizes; plus slivers (kind 3), zero-area triangles (4) and normals whose components sum to zero (5: norm.a[3] = 0,
    crixus_d.cu:1099-1101).  Kinds 1, 6 (u+v = 1+eps can be hit exactly) and 3 are rounding-sensitive: their outcome depends on
    the compiler's multiply-add fusion, so only the reference CUDA binary (golden fixture) pins them."""
    rng = np.random.default_rng(1000 + seed)
    tris = np.asarray(tris, dtype=np.float64)
    bary = tris.mean(1)
    nbe = bary.shape[0]
    ext = float((tris.reshape(-1, 3).max(0) - tris.reshape(-1, 3).min(0)).max())
    eps = 1e-5 * ext
    out = []
    for k in range(n):
        b = bary[rng.integers(nbe)]
        kind = k % 8
        size = ext * 10 ** rng.uniform(-1.7, 0.5)
        e1 = rng.normal(size=3)
        e1 /= np.linalg.norm(e1)
        e2 = rng.normal(size=3)
        e2 -= e2.dot(e1) * e1
        e2 /= np.linalg.norm(e2)
        if kind == 5:                      # normal (1,-1,0)/sqrt2: components sum to zero
            e1, e2 = np.array([1.0, 1.0, 0.0]) / np.sqrt(2), np.array([0.0, 0.0, 1.0])
        if kind == 6:                      # axis-aligned, in the plane of a wall
            e1, e2 = np.array([1.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0])
        ba, ca = size * e1, size * rng.uniform(0.3, 1.5) * (0.3 * e1 + e2) if kind != 3 else size * (e1 + 1e-4 * e2)   # 3: sliver
        if kind == 4:
            ca = 2.0 * ba                  # zero area
        u, v = [(0.0, 0.0), (0.5, 0.5), (0.3, 0.0), (1 / 3, 1 / 3), (0.2, 0.2), (0.25, 0.25), (0.0, 1.0), (0.6, 0.1)][kind]
        du, dv = rng.choice([-1.5, -0.5, 0.0, 0.5, 1.5], size=2) * eps
        lift = rng.choice([0.0, 0.0, 0.3, 0.9, 1.1, 3.0]) * eps * np.cross(e1, e2)
        v0 = b - (u + du) * ba - (v + dv) * ca - lift
        if kind not in skip:
            out.append([v0, v0 + ba, v0 + ca])
    return np.ascontiguousarray(np.array(out), dtype=np.float32)


def generated(seed):
    """A generated parity case: jittered closed tank (box + geometry fill, free-surface lid at a random height for even seeds)
    or, for seed % 3 == 2, an x/y-periodic channel; dr, dr_wall, jitter amplitude and the fill seed are drawn from the seed."""
    rng = np.random.default_rng(500 + seed)
    dr = float(rng.choice([0.07, 0.08, 0.1]))
    if seed % 3 == 2:
        v, t, n = mg.channel(n=int(rng.integers(10, 18)), dr=dr, height=8 * 1.25 * dr)
        L = float(v[:, 0].max())
        c = dict(swap_normals=True, periodic=(True, True, False), container=None, fshape=None,
                 fills=[("geometry", (L * rng.uniform(0.2, 0.8), L * rng.uniform(0.2, 0.8), 4 * 1.25 * dr), dr * float(rng.choice([0.5, 1.0, 1.5])))])
    else:
        v, t, n = mg.box_tank((0, 0, 0), (1, 1, 1), 0.1, inward=True)
        interior = ((v > 1e-6) & (v < 1 - 1e-6)).sum(1) == 2
        v = v.copy()
        ids = np.nonzero(interior)[0]
        ids = ids[np.lexsort((v[ids, 2], v[ids, 1], v[ids, 0]))]
        v[ids] += rng.uniform(-1, 1, size=(ids.size, 3)).astype(np.float32) * np.float32(rng.choice([0.0, 0.002, 0.004]))
        n = mg.geometric_normals(v, t)
        n[(n * (v[t].mean(1) - 0.5)).sum(1) > 0] *= -1
        zl = float(rng.uniform(0.3, 0.8))
        fs_t = np.array([[[0, 0, zl], [1, 0, zl], [1, 1, zl]], [[0, 0, zl], [1, 1, zl], [0, 1, zl]]], dtype=np.float32)
        fs_n = np.array([[0, 0, 1], [0, 0, 1]], dtype=np.float32)
        c = dict(swap_normals=False, periodic=(False, False, False), container=None,
                 fills=[("box", (0.3, 0.3, 0.1), (0.5, 0.6, 0.2)),
                        ("geometry", tuple(float(x) for x in rng.uniform(0.15, 0.85, 2)) + (zl * 0.5,), dr * float(rng.choice([0.5, 1.0, 1.5])))],
                 fshape=(fs_t, fs_n) if seed % 2 == 0 else None)
    c.update(name="generated%d" % seed, tris=np.ascontiguousarray(v[t], dtype=np.float32), normals=np.ascontiguousarray(n, dtype=np.float32),
             dr=np.float32(dr), zero_open=False, special=[])
    return c
