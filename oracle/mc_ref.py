"""oracle/mc_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Marching cubes for the meshing step that follows the dense SDF query (src/utils/Mesher.py:230-258):

    verts, faces, _, _ = skimage.measure.marching_cubes(volume[x,y,z], level, spacing=(dx,dy,dz));  verts += (x0,y0,z0)

PARITY UNPINNED for the triangulation: scikit-image (the reference's dependency, `requirements.txt`) is absent from this image
and from /root/reference, and its Lewiner case tables are not something to restate from memory.  What IS restated, and what
the CUDA path is held to bit for bit, is a complete marching cubes of the classic kind:

  * a vertex on every grid edge whose end values straddle the level (inside = value < level), at the linear interpolation
    t = (level - v0) / (v1 - v0) -- the same vertex set and positions as any marching-cubes implementation;
  * a 256-case triangle table GENERATED here from first principles (build_tables): on each cube face the crossing points are
    joined into directed segments (entry -> exit in counter-clockwise order seen from outside; on an ambiguous face the two
    inside corners are cut off separately -- a rule that only looks at the face's own four signs, so the two cubes sharing a
    face always agree and the mesh is crack-free), the segments chain into closed loops, each loop is fanned into triangles.
    It differs from Lewiner's tables only in how ambiguous configurations are triangulated.

Checks that pin this restatement (tests/test_mesh.py): closed oriented 2-manifold on analytic SDFs (every directed edge once
in each direction), Euler characteristic 2 for a sphere, vertices within O(h^2) of the analytic surface, complement symmetry
of the table.

Conventions: cube corner c has offsets (c & 1, (c >> 1) & 1, (c >> 2) & 1) along (x, y, z).  Edge e = 4 * axis + (bu + 2 * bv)
runs along `axis` from the corner whose other two coordinates are (bu, bv) in the order (y,z) / (x,z) / (x,y) for axis x / y / z.
Vertices are numbered by owner grid point in the volume's memory order, then by axis; faces by cell in the same order, then by
the table's triangle order.  The volume is addressed as vol[iy, ix, iz] -- the layout the dense query writes
(torch.meshgrid(indexing='xy') flattened, Mesher.py:192-193).
"""
import numpy as np

AXIS_OTHER = ((1, 2), (0, 2), (0, 1))            # (u, v) axes of an edge along axis a


def corner_offset(c):
    return (c & 1, (c >> 1) & 1, (c >> 2) & 1)


def edge_corners(e):
    a, k = divmod(e, 4)
    u, v = AXIS_OTHER[a]
    off = [0, 0, 0]
    off[u] = k & 1; off[v] = k >> 1
    c0 = off[0] | (off[1] << 1) | (off[2] << 2)
    return c0, c0 | (1 << a)


def edge_of(c0, c1):
    a = (c0 ^ c1).bit_length() - 1
    lo = min(c0, c1)
    o = corner_offset(lo)
    u, v = AXIS_OTHER[a]
    return 4 * a + (o[u] + 2 * o[v])


def _faces():
    """6 faces as corner 4-cycles, counter-clockwise seen from OUTSIDE the cube."""
    out = []
    for a in range(3):
        u, v = (a + 1) % 3, (a + 2) % 3                       # (a, u, v) is a cyclic permutation of (x, y, z)
        for side in (0, 1):
            cyc = []
            for (bu, bv) in ((0, 0), (1, 0), (1, 1), (0, 1)):
                o = [0, 0, 0]
                o[a] = side; o[u] = bu; o[v] = bv
                cyc.append(o[0] | (o[1] << 1) | (o[2] << 2))
            out.append(cyc if side == 1 else cyc[::-1])
    return out


def build_tables():
    """tri_table[config] = list of (e0, e1, e2) edge triples; config bit c set <=> corner c inside (value < level).
    Triangle normals (right-hand rule) point from inside to outside, i.e. towards larger values."""
    faces = _faces()
    table = []
    for cfg in range(256):
        inside = [(cfg >> c) & 1 for c in range(8)]
        nxt = {}                                               # directed segments between edge-crossing vertices
        for cyc in faces:
            entries, exits = {}, {}
            for k in range(4):
                p, q = cyc[k], cyc[(k + 1) % 4]
                if inside[p] != inside[q]:
                    e = edge_of(p, q)
                    if inside[q]:
                        entries[q] = e                         # outside -> inside: the crossing just BEFORE inside corner q
                    else:
                        exits[p] = e                           # inside -> outside: the crossing just AFTER inside corner p
            if len(entries) == 1:
                (ein,), (eout,) = entries.values(), exits.values()
                nxt[ein] = eout
            elif len(entries) == 2:                            # ambiguous face: cut the two inside corners off separately
                for p in entries:
                    nxt[entries[p]] = exits[p]
        tris, seen = [], set()
        for start in sorted(nxt):
            if start in seen:
                continue
            loop, e = [], start
            while e not in seen:
                seen.add(e); loop.append(e); e = nxt[e]
            assert e == start and len(loop) >= 3
            tris += _triangulate(loop)
        table.append(tris)
    return table


def _edge_faces(e):
    """The two cube faces (axis, side) an edge lies on."""
    a, k = divmod(e, 4)
    u, v = AXIS_OTHER[a]
    return {(u, k & 1), (v, k >> 1)}


def _triangulate(loop):
    """Triangles of one closed loop of crossing vertices (normals towards larger values).  A diagonal joining two vertices
    of the SAME cube face would lie in that face, where the neighbouring cube may draw the same diagonal: four triangles on
    one edge.  So: the first triangulation (fans first, then all others) without an in-face diagonal."""
    n = len(loop)

    def ok(i, j):                                  # is the chord loop[i]-loop[j] admissible (polygon side, or not in a face)?
        if (i - j) % n in (1, n - 1):
            return True
        return not (_edge_faces(loop[i]) & _edge_faces(loop[j]))

    def rec(idx):                                  # all triangulations of the sub-polygon idx (list of loop positions)
        if len(idx) < 3:
            return [[]]
        if len(idx) == 3:
            return [[tuple(idx)]] if ok(idx[0], idx[2]) and ok(idx[0], idx[1]) and ok(idx[1], idx[2]) else []
        out = []
        a, b = idx[0], idx[-1]
        if not ok(a, b):
            return []
        for m in range(1, len(idx) - 1):
            c = idx[m]
            if not (ok(a, c) and ok(c, b)):
                continue
            for left in rec(idx[:m + 1]):
                for right in rec(idx[m:]):
                    out.append(left + [(a, c, b)] + right)
        return out

    for rot in range(n):                           # fans first: apex = loop[rot]
        fan = [(rot, (rot + k) % n, (rot + k + 1) % n) for k in range(1, n - 1)]
        if all(ok(t[0], t[1]) and ok(t[1], t[2]) and ok(t[0], t[2]) for t in fan):
            return [(loop[i], loop[j], loop[k]) for i, j, k in fan]
    alls = rec(list(range(n)))
    assert alls, f"no admissible triangulation for loop {loop}"
    return [(loop[i], loop[j], loop[k]) for i, j, k in alls[0]]


_TABLE = None


def tri_table():
    global _TABLE
    if _TABLE is None:
        _TABLE = build_tables()
    return _TABLE


def max_triangles():
    return max(len(t) for t in tri_table())


def edge_owner(e):
    """(dx, dy, dz, axis): the grid point (relative to the cell origin) that owns edge e and the axis it runs along."""
    a, k = divmod(e, 4)
    u, v = AXIS_OTHER[a]
    off = [0, 0, 0]
    off[u] = k & 1; off[v] = k >> 1
    return off[0], off[1], off[2], a


def marching_cubes(vol_yxz, level, origin, spacing):
    """vol_yxz[iy, ix, iz] fp32.  Returns verts (V,3) fp32 = origin + spacing * (index + t) evaluated in fp32, faces (F,3) int64,
    and keys (V,) int64 = 3 * flat point index + axis (what a multi-GPU merge welds seam vertices by)."""
    vol = np.ascontiguousarray(vol_yxz, dtype=np.float32)
    ny, nx, nz = vol.shape
    lvl = np.float32(level)
    ins = vol < lvl
    org = np.asarray(origin, dtype=np.float32); sp = np.asarray(spacing, dtype=np.float32)
    # crossing flags per point and axis (axis 0 = x = array axis 1, axis 1 = y = array axis 0, axis 2 = z = array axis 2)
    flags = np.zeros((ny, nx, nz, 3), dtype=bool)
    flags[:, :-1, :, 0] = ins[:, :-1, :] != ins[:, 1:, :]
    flags[:-1, :, :, 1] = ins[:-1, :, :] != ins[1:, :, :]
    flags[:, :, :-1, 2] = ins[:, :, :-1] != ins[:, :, 1:]
    vid = np.cumsum(flags.reshape(-1)) - 1                      # vertex index of (point, axis), point-major / axis-minor
    vid = vid.reshape(ny, nx, nz, 3)
    iy, ix, iz, ax = np.nonzero(flags)
    v0 = vol[iy, ix, iz]
    v1 = vol[iy + (ax == 1), ix + (ax == 0), iz + (ax == 2)]
    t = ((lvl - v0) / (v1 - v0)).astype(np.float32)
    idx = np.stack([ix, iy, iz], axis=1).astype(np.float32)
    idx[np.arange(len(ax)), ax] += t
    verts = (org[None, :] + sp[None, :] * idx).astype(np.float32)
    keys = ((iy.astype(np.int64) * nx + ix) * nz + iz) * 3 + ax
    # cells
    cfgs = np.zeros((ny - 1, nx - 1, nz - 1), dtype=np.int32)
    for c in range(8):
        ox, oy, oz = corner_offset(c)
        cfgs |= ins[oy:ny - 1 + oy, ox:nx - 1 + ox, oz:nz - 1 + oz].astype(np.int32) << c
    table = tri_table()
    faces = []
    cy, cx, cz = np.nonzero((cfgs != 0) & (cfgs != 255))
    for y, x, z in zip(cy, cx, cz):
        for tri in table[int(cfgs[y, x, z])]:
            f = []
            for e in tri:
                dx, dy, dz, a = edge_owner(e)
                f.append(vid[y + dy, x + dx, z + dz, a])
            faces.append(f)
    faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    return verts, faces, keys


def mesh_stats(verts, faces):
    """Directed-edge bookkeeping: (closed_manifold, euler_characteristic)."""
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
    fwd = e[:, 0].astype(np.int64) * (len(verts) + 1) + e[:, 1]
    bwd = e[:, 1].astype(np.int64) * (len(verts) + 1) + e[:, 0]
    uf, cf = np.unique(fwd, return_counts=True)
    closed = bool((cf == 1).all() and np.array_equal(uf, np.unique(bwd)))
    n_edges = len(np.unique(np.minimum(fwd, bwd)))
    used = len(np.unique(faces))
    return closed, used - n_edges + len(faces)
