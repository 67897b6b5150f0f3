"""CPU oracle of the mesh extraction -- TEST INFRASTRUCTURE ONLY, like the rest of oracle/.

PARITY UNPINNED: the reference meshes on the CPU with KrisLibrary's Geometry::SparseTSDFReconstruction::ExtractMesh
(examples/ros_camera_driver/ros_offline.cc:258-318); KrisLibrary (Klampt) is an external dependency that is neither
vendored in /root/reference nor installed here, so there is no reference output to pin against.  What this oracle
states is the engine's own definition of the mesh, applied DIRECTLY (it does not read the engine's triangle table):

  * voxel centres sit at (grid + 0.5) * voxel_size -- the `+ ofs` of ros_offline.cc:281-284;
  * a cell is the cube between 8 neighbouring voxel centres, based at a voxel of a selected block (selection =
    GatherVoxels' rule, voxel_tsdf.cu:14-32 / voxel_tsdf.cuh:21-26; all blocks when bbox is None);
  * a cell is meshed only if all 8 corner voxels are allocated AND observed (weight > 0): never-observed voxels carry the
    pool's initial tsdf = -1 (voxel_mem.cu:48-50) and would otherwise put a spurious sheet between observed free
    space and unobserved space;
  * inside = tsdf < 0; the cube is split into the 6 Kuhn tetrahedra around the diagonal 0-7; per tetrahedron one
    triangle (1 or 3 corners inside) or the quad ac, ad, bd, bc as (ac, ad, bd) + (ac, bd, bc) (2 inside: a < b, outside
    c < d in the tetrahedron's corner order); normals point from inside to outside;
  * a vertex on the edge p < q (cube corner indices) is P_p + t * (P_q - P_p), t = f_p / (f_p - f_q), all float32.
Triangles are returned in no particular order: compare with canonical_triangles().
"""
import numpy as np

TETS = [(0, 1, 3, 7), (0, 3, 2, 7), (0, 2, 6, 7), (0, 6, 4, 7), (0, 4, 5, 7), (0, 5, 1, 7)]
CORNER = np.array([[c & 1, (c >> 1) & 1, (c >> 2) & 1] for c in range(8)], np.int32)


def select_blocks(keys, voxel_size, bbox):
    """GatherVoxels' block selection (voxel_tsdf.cu:14-32): block fully inside the inclusive short-voxel bound."""
    keys = np.asarray(keys, np.int32)
    if bbox is None:
        return np.ones(len(keys), bool)
    scale = np.float32(1.0 / float(np.float32(voxel_size)))
    g = np.clip(np.trunc(np.asarray(bbox, np.float32) * scale), -32768, 32767).astype(np.int32)
    base = ((keys << 3).astype(np.int16)).astype(np.int32)
    return ((base[:, 0] >= g[0]) & (base[:, 1] >= g[2]) & (base[:, 2] >= g[4]) & (base[:, 0] + 7 <= g[1]) &
            (base[:, 1] + 7 <= g[3]) & (base[:, 2] + 7 <= g[5]))


_CASES = {}


def cell_triangles(mask):
    """Oriented triangles of one cell as edge triples ((p, q), ...) for the 8-bit inside mask -- the rule of the module
    docstring applied directly (cached per mask)."""
    if mask in _CASES:
        return _CASES[mask]
    pos = CORNER.astype(np.float64)
    tris = []
    for tet in TETS:
        tin = [c for c in tet if (mask >> c) & 1]
        tout = [c for c in tet if not (mask >> c) & 1]
        if len(tin) in (0, 4):
            continue
        if len(tin) == 1:
            cand = [[(tin[0], o) for o in tout]]
        elif len(tin) == 3:
            cand = [[(i, tout[0]) for i in tin]]
        else:
            (a, b), (c, d) = tin, tout
            cand = [[(a, c), (a, d), (b, d)], [(a, c), (b, d), (b, c)]]
        towards_outside = pos[tout].mean(0) - pos[tin].mean(0)
        for tri in cand:
            tri = [(min(e), max(e)) for e in tri]
            m = [0.5 * (pos[e[0]] + pos[e[1]]) for e in tri]
            if np.dot(np.cross(m[1] - m[0], m[2] - m[0]), towards_outside) < 0:
                tri = [tri[0], tri[2], tri[1]]
            tris.append(tuple(tri))
    _CASES[mask] = tris
    return tris


def extract_mesh(keys, tsdf, rgbw, voxel_size, bbox=None):
    """keys int16[n,3], tsdf float32[n,512], rgbw uint8[n,512,4] (Oracle.export()).  Returns float32[m, 3, 3]."""
    vs = np.float32(voxel_size)
    keys = np.asarray(keys, np.int32)
    index = {tuple(k): i for i, k in enumerate(keys.tolist())}
    sel = select_blocks(keys, voxel_size, bbox)
    lx, ly, lz = np.meshgrid(np.arange(8), np.arange(8), np.arange(8), indexing="ij")  # cell base voxel in the block
    cells_f, cells_g = [], []
    for bi in np.nonzero(sel)[0]:
        bx, by, bz = keys[bi].tolist()
        f = np.ones((9, 9, 9), np.float32)     # 9^3 halo of tsdf / observed; absent blocks = unobserved
        obs = np.zeros((9, 9, 9), bool)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    # neighbour block coordinates wrap like the reference's short arithmetic
                    j = index.get(tuple(np.array([bx + dx, by + dy, bz + dz]).astype(np.int16).tolist()))
                    if j is None:
                        continue
                    t = tsdf[j].reshape(8, 8, 8).transpose(2, 1, 0)  # [x, y, z] from index x + 8y + 64z
                    w = rgbw[j][:, 3].reshape(8, 8, 8).transpose(2, 1, 0) > 0
                    sx, sy, sz = (slice(0, 8) if d == 0 else slice(8, 9) for d in (dx, dy, dz))
                    qx, qy, qz = (slice(0, 8) if d == 0 else slice(0, 1) for d in (dx, dy, dz))
                    f[sx, sy, sz] = t[qx, qy, qz]
                    obs[sx, sy, sz] = w[qx, qy, qz]
        fc = np.stack([f[lx + c[0], ly + c[1], lz + c[2]] for c in CORNER], -1)       # [8, 8, 8, corner]
        oc = np.stack([obs[lx + c[0], ly + c[1], lz + c[2]] for c in CORNER], -1)
        inside = fc < 0
        mixed = oc.all(-1) & inside.any(-1) & ~inside.all(-1)
        if not mixed.any():
            continue
        cx, cy, cz = np.nonzero(mixed)
        cells_f.append(fc[cx, cy, cz])
        cells_g.append(np.stack([(bx << 3) + cx, (by << 3) + cy, (bz << 3) + cz], -1))
    if not cells_f:
        return np.zeros((0, 3, 3), np.float32)
    F = np.concatenate(cells_f)                       # [m, 8]
    G = np.concatenate(cells_g).astype(np.int32)      # [m, 3] voxel coordinates of corner 0
    masks = ((F < 0) * (1 << np.arange(8))).sum(-1)
    # voxel centres of the 8 corners, float32, with the reference's short wrap of the coordinates
    P = ((G[:, None, :] + CORNER[None]).astype(np.int16).astype(np.float32) + np.float32(0.5)) * vs   # [m, 8, 3]
    out = []
    for mask in np.unique(masks).tolist():
        rows = np.nonzero(masks == mask)[0]
        f, pos = F[rows], P[rows]
        for tri in cell_triangles(mask):
            v = []
            for (p, q) in tri:
                t = f[:, p] / (f[:, p] - f[:, q])
                v.append(pos[:, p] + t[:, None] * (pos[:, q] - pos[:, p]))
            out.append(np.stack(v, 1))
    return np.concatenate(out).astype(np.float32)


def canonical_triangles(tris):
    """Rotate every triangle so that its lexicographically smallest vertex comes first (orientation kept), then sort the
    triangles: two meshes are the same set of oriented triangles iff these arrays are equal."""
    t = np.asarray(tris, np.float32).reshape(-1, 3, 3)
    if len(t) == 0:
        return t
    key = t.view(np.uint32).astype(np.int64)  # order by bit pattern: exact, NaN-free inputs
    first = np.lexsort((key[:, :, 2], key[:, :, 1], key[:, :, 0]), axis=1)[:, 0]
    idx = (first[:, None] + np.arange(3)[None]) % 3
    t = np.take_along_axis(t, idx[:, :, None], 1)
    flat = t.reshape(len(t), 9).view(np.uint32)
    return t[np.lexsort(flat.T[::-1])]


def mesh_properties(tris):
    """Size-independent checks: signed volume, area, and how many undirected edges are used by exactly 2 triangles
    (every edge, for a closed surface) with opposite directions (consistent orientation)."""
    t = np.asarray(tris, np.float64).reshape(-1, 3, 3)
    area2 = np.linalg.norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]), axis=1)
    t = t[area2 > 0]  # drop degenerate triangles (a crossing exactly at a voxel centre)
    vol = float(np.einsum("ij,ij->i", t[:, 0], np.cross(t[:, 1], t[:, 2])).sum() / 6.0)
    verts, inv = np.unique(np.asarray(tris, np.float32).reshape(-1, 3, 3)[area2 > 0].reshape(-1, 3).view(np.uint32), axis=0, return_inverse=True)
    tri_idx = inv.reshape(-1, 3)
    e = np.concatenate([tri_idx[:, [0, 1]], tri_idx[:, [1, 2]], tri_idx[:, [2, 0]]])
    und = np.sort(e, 1)
    uniq, cnt = np.unique(und, axis=0, return_counts=True)
    # directed balance: each undirected edge should be traversed once in each direction
    sign = np.where(e[:, 0] < e[:, 1], 1, -1)
    _, inv_e = np.unique(und, axis=0, return_inverse=True)
    balance = np.bincount(inv_e.ravel(), weights=sign, minlength=len(uniq))
    return {"triangles": int(len(t)), "volume": vol, "area": float(area2[area2 > 0].sum() / 2.0), "edges": int(len(uniq)),
            "edges_shared_by_2": int((cnt == 2).sum()), "edges_unbalanced": int((balance != 0).sum()), "vertices": int(len(verts))}


def sphere_volume(radius=0.9, vs=0.05, half=3, observed=None):
    """Hand-built test volume: (2 half)^3 blocks holding the truncated SDF of a sphere (slightly off-centre), every voxel
    observed unless `observed(centres) -> bool mask` says otherwise.  Returns keys, tsdf, rgbw, voxel_size."""
    rng = range(-half, half)
    keys = np.array([[x, y, z] for z in rng for y in rng for x in rng], np.int16)
    k = np.arange(512)
    lx, ly, lz = k & 7, (k >> 3) & 7, k >> 6
    tsdf = np.zeros((len(keys), 512), np.float32)
    rgbw = np.zeros((len(keys), 512, 4), np.uint8)
    rgbw[:, :, 3] = 5
    for i, (bx, by, bz) in enumerate(keys.tolist()):
        c = (np.stack([bx * 8 + lx, by * 8 + ly, bz * 8 + lz], -1) + 0.5) * vs
        tsdf[i] = np.clip((np.linalg.norm(c - np.array([0.013, -0.021, 0.007]), axis=1) - radius) / 0.3, -1, 1)
        if observed is not None:
            rgbw[i, ~observed(c), 3] = 0
    return keys, tsdf, rgbw, vs
