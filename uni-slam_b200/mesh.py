"""Mesh extraction on the device: what src/utils/Mesher.py:197-276 (get_mesh) does after the dense SDF
query -- marching cubes on the volume, vertex colours from the colour field, vertices / scale, a PLY file -- without
moving the 504 MiB volume to the host.

  volume (a y-slab as steps.DenseSdfQuery writes it, vol[(iy*nx + ix)*nz + iz])
    -> usl_mc_classify -> usl_scan_u8 x2 -> usl_mc_emit            (csrc/mesh.cu; indexed triangles, shared vertices)
    -> vertex colours: the fused field query on the vertex positions (eval_points(...)[..., :3], Mesher.py:259-267)
    -> weld the seam vertices of neighbouring slabs by their global edge key (multi-GPU / multi-slab)
    -> binary little-endian PLY (vertex x y z red green blue, face list) -- the format trimesh.export writes

Triangulation: a generated, crack-free 256-case table (csrc/mc_tables.h); the reference's skimage uses Lewiner's tables,
which differ in ambiguous configurations only (oracle/mc_ref.py, parity unpinned for the triangulation -- scikit-image is
not in this image).

Culling (src/tools/cull_mesh.py), on the device arrays marching cubes left:
  MeshCuller.cull_by_frames  = cull_mesh (cull_mesh.py:31-109; Mapper.py:556,570): drop what no frame sees
  MeshCuller.cull_by_hull    = cull_out_bound_mesh (cull_mesh.py:112-148; Mesher.py:274): drop what lies outside the convex bound
The convex bound itself (Mesher.get_bound_from_frames: open3d TSDF fusion + convex hull, Mesher.py:64-131) is an INPUT here
(hull vertices + faces); building it is third-party geometry outside the path.
"""
from ctypes import byref, c_int64
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import ops
from ._lib import call, cptr, ptr, stream


class MeshExtractor:
    """Marching cubes on one y-slab of the SDF volume.  axes = the three coordinate arrays of Mesher.get_grid_uniform
    (np.linspace, fp32); spacing = axis[2] - axis[1] as the reference passes it to skimage (Mesher.py:241-243)."""

    def __init__(self, axes: Sequence[torch.Tensor], level: float = 0.0):
        ax = [a.detach().float().cpu() for a in axes]
        self.nx, self.ny, self.nz = [int(a.numel()) for a in ax]
        self.origin = [float(a[0]) for a in ax]
        self.spacing = [float(a[2] - a[1]) if a.numel() > 2 else float(a[-1] - a[0]) for a in ax]
        self.level = float(level)

    def run(self, vol: torch.Tensor, y_begin: int = 0, y_end: Optional[int] = None, halo: bool = False, keys: bool = False):
        """vol: fp32 device tensor of (rows, nx, nz) values, rows = y_end - y_begin (+1 halo row when halo=True, i.e. the slab
        is not the volume's last one).  Returns verts (V,3) fp32, faces (T,3) int32 (device) [, keys (V,) int64]."""
        y_end = self.ny if y_end is None else y_end
        own = y_end - y_begin
        rows = own + (1 if halo else 0)
        n = rows * self.nx * self.nz
        if vol.numel() != n or vol.dtype != torch.float32 or not vol.is_cuda or not vol.is_contiguous():
            raise ValueError(f"MeshExtractor: volume slab must be a contiguous fp32 CUDA tensor of {rows}x{self.nx}x{self.nz} values")
        dev = vol.device
        a = L.McArgs()
        a.vol, a.nx, a.nz, a.rows, a.own_rows, a.y_begin, a.level = ptr(vol), self.nx, self.nz, rows, own, y_begin, self.level
        for d in range(3):
            a.origin[d] = self.origin[d]; a.spacing[d] = self.spacing[d]
        pflags = torch.empty((n,), device=dev, dtype=torch.uint8); ctri = torch.empty((n,), device=dev, dtype=torch.uint8)
        a.pflags, a.ctri = ptr(pflags), ptr(ctri)
        st = stream()
        call("usl_mc_classify", byref(a), st)
        nb = c_int64(0)
        call("usl_scan_u8_blocks", n, byref(nb))
        sums = torch.empty((max(nb.value, 1),), device=dev, dtype=torch.int32)
        voff = torch.empty((n,), device=dev, dtype=torch.int32); toff = torch.empty((n,), device=dev, dtype=torch.int32)
        totals = torch.zeros((2,), device=dev, dtype=torch.int32)
        call("usl_scan_u8", ptr(pflags), n, 1, ptr(voff), ptr(sums), ptr(totals[0:1]), st)
        call("usl_scan_u8", ptr(ctri), n, 0, ptr(toff), ptr(sums), ptr(totals[1:2]), st)
        V, T = [int(v) for v in totals.cpu()]                      # the one host read: the output sizes
        verts = torch.empty((V, 3), device=dev, dtype=torch.float32)
        faces = torch.empty((T, 3), device=dev, dtype=torch.int32)
        vkeys = torch.empty((V,), device=dev, dtype=torch.int64) if keys else None
        if V > 0:
            a.voff, a.toff, a.verts, a.faces, a.vkeys = ptr(voff), ptr(toff), ptr(verts), ptr(faces), ptr(vkeys)
            call("usl_mc_emit", byref(a), st)
        return (verts, faces, vkeys) if keys else (verts, faces)


def vertex_colors(meta: ops.FieldMeta, sdf_table, rgb_table, dec, verts: torch.Tensor, bound: torch.Tensor) -> torch.Tensor:
    """eval_points(verts)[..., :3] (Mesher.py:134-166, 259-267): normalise by the scene bound, clamp, both grids + decoders
    (the fused forward), colour channels; uint8 after clip(0,1)*255 as trimesh stores vertex colours."""
    if verts.shape[0] == 0:
        return torch.empty((0, 3), device=verts.device, dtype=torch.uint8)
    b = bound.to(verts.device, torch.float32)
    x = ((verts - b[:, 0]) / (b[:, 1] - b[:, 0])).contiguous()
    with torch.no_grad():
        raw = ops.field_points(meta, x, sdf_table.detach(), rgb_table.detach(), [d.detach() for d in dec])
    return (raw[:, :3].clamp(0, 1) * 255).round().to(torch.uint8)


def hull_planes(hull_verts, hull_faces) -> np.ndarray:
    """(F,4) fp32 outward planes of a closed convex hull given as vertices + triangles (what trimesh.Trimesh holds for
    Mesher.get_bound_from_frames' return value): unit normal n and offset d with n.x + d <= 0 inside; the orientation is
    fixed by the hull's vertex centroid, so the winding of the faces does not matter."""
    hv = np.asarray(hull_verts, dtype=np.float64); hf = np.asarray(hull_faces, dtype=np.int64)
    a, b, c = hv[hf[:, 0]], hv[hf[:, 1]], hv[hf[:, 2]]
    n = np.cross(b - a, c - a)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    ok = ln[:, 0] > 0                                                # degenerate (zero-area) hull faces carry no constraint
    n, a = n[ok] / ln[ok], a[ok]
    d = -(n * a).sum(axis=1)
    flip = (n @ hv.mean(axis=0) + d) > 0
    n[flip] *= -1; d[flip] *= -1
    return np.concatenate([n, d[:, None]], axis=1).astype(np.float32)


class MeshCuller:
    """cull_mesh / cull_out_bound_mesh (src/tools/cull_mesh.py) on device-resident meshes: verts (V,3) fp32, faces (T,3)
    int32, optional colours (V,3) uint8 -- the arrays MeshExtractor.run / vertex_colors return.  Every method returns the
    culled (verts, faces, colors) as new device tensors, vertex and face order preserved (trimesh's update_faces +
    remove_unreferenced_vertices); the one host read per call is the pair of output sizes."""

    def __init__(self, H, W, fx, fy, cx, cy, truncation):
        self.cam = (int(H), int(W), float(fx), float(fy), float(cx), float(cy))
        self.truncation = float(truncation)
        self.frames_per_cta = 0           # 0 = library default (16 depth frames in flight, L2-resident)

    # ---- per-vertex tests ----
    def seen_by_frames(self, verts: torch.Tensor, c2ws: torch.Tensor, depths: Optional[torch.Tensor], eval_rec: bool,
                       seen: Optional[torch.Tensor] = None, w2c: Optional[torch.Tensor] = None) -> torch.Tensor:
        """NOT whole_mask of cull_mesh.py:58-99, (V,) uint8.  c2ws (K,4,4): estimate_c2w_list[:idx+1] (or the ground-truth
        poses); depths (K,H,W) fp32 sensor depth (required with eval_rec = cfg['meshing']['eval_rec']).  Pass `seen` to
        accumulate over several calls (ranges of frames); pass `w2c` (K,4,4) = torch.inverse(c2ws) when it is already at hand."""
        H, W, fx, fy, cx, cy = self.cam
        V, K = verts.shape[0], c2ws.shape[0]
        if seen is None:
            seen = torch.zeros((V,), device=verts.device, dtype=torch.uint8)
        a = L.CullFramesArgs()
        a.verts, a.V = cptr(verts, torch.float32, V * 3, "verts (V,3)"), V
        if w2c is None:
            w2c = torch.inverse(c2ws.to(verts.device, torch.float32)).contiguous()      # cull_mesh.py:69, batched
        a.w2c = cptr(w2c, torch.float32, K * 16, "c2ws (K,4,4)")
        if eval_rec and depths is None:
            raise ValueError("MeshCuller: eval_rec needs the depth frames (K,H,W)")
        a.depths = cptr(depths, torch.float32, K * H * W, "depths (K,H,W)") if eval_rec else None
        a.K, a.H, a.W, a.fx, a.fy, a.cx, a.cy = K, H, W, fx, fy, cx, cy
        a.truncation, a.eval_rec, a.frames_per_cta = self.truncation, 1 if eval_rec else 0, self.frames_per_cta
        a.seen = cptr(seen, torch.uint8, V, "seen (V,)")
        call("usl_mesh_cull_frames", byref(a), stream())
        return seen

    @staticmethod
    def inside_hull(verts: torch.Tensor, planes) -> torch.Tensor:
        """mesh_bound.contains(vertices) for a convex bound given as outward planes (F,4) (see hull_planes), (V,) uint8."""
        V = verts.shape[0]
        pl = torch.as_tensor(planes, dtype=torch.float32).to(verts.device).contiguous()
        inside = torch.empty((V,), device=verts.device, dtype=torch.uint8)
        call("usl_mesh_cull_hull", cptr(verts, torch.float32, V * 3, "verts (V,3)"), V, ptr(pl), pl.shape[0], ptr(inside), stream())
        return inside

    # ---- face rule + compaction ----
    @staticmethod
    def filter_faces(verts, faces, colors, vmask, require_all: bool, vertex_kept: Optional[list] = None):
        """vertex_kept: pass a list to receive the (V,) uint8 mask of the vertices that survive -- any other per-vertex array
        (the slab's edge keys for mesh.weld) is compacted with it: keys[mask.bool()].  Multi-GPU: every rank culls its own
        slab's mesh (a vertex's mark depends on its position only, so the two copies of a seam vertex agree), then weld."""
        V, T = verts.shape[0], faces.shape[0]
        dev = verts.device
        if V == 0 or T == 0:                                      # nothing can be referenced: the empty mesh
            if vertex_kept is not None:
                vertex_kept.append(torch.zeros((V,), device=dev, dtype=torch.uint8))
            return (torch.empty((0, 3), device=dev, dtype=torch.float32), torch.empty((0, 3), device=dev, dtype=torch.int32),
                    torch.empty((0, 3), device=dev, dtype=torch.uint8) if colors is not None else None)
        st = stream()
        keep = torch.empty((T,), device=dev, dtype=torch.uint8); vref = torch.zeros((V,), device=dev, dtype=torch.uint8)
        call("usl_mesh_face_keep", cptr(faces, torch.int32, T * 3, "faces (T,3)"), T, cptr(vmask, torch.uint8, V, "vertex mask (V,)"), V,
             1 if require_all else 0, ptr(keep), ptr(vref), st)
        nb = c_int64(0)
        call("usl_scan_u8_blocks", max(V, T), byref(nb))
        sums = torch.empty((max(nb.value, 1),), device=dev, dtype=torch.int32)
        voff = torch.empty((V,), device=dev, dtype=torch.int32); foff = torch.empty((T,), device=dev, dtype=torch.int32)
        totals = torch.zeros((2,), device=dev, dtype=torch.int32)
        call("usl_scan_u8", ptr(vref), V, 0, ptr(voff), ptr(sums), ptr(totals[0:1]), st)
        call("usl_scan_u8", ptr(keep), T, 0, ptr(foff), ptr(sums), ptr(totals[1:2]), st)
        V2, T2 = [int(v) for v in totals.cpu()]
        if vertex_kept is not None:
            vertex_kept.append(vref)
        verts_out = torch.empty((V2, 3), device=dev, dtype=torch.float32); faces_out = torch.empty((T2, 3), device=dev, dtype=torch.int32)
        colors_out = torch.empty((V2, 3), device=dev, dtype=torch.uint8) if colors is not None else None
        if V2 > 0:
            call("usl_mesh_compact", ptr(verts), cptr(colors, torch.uint8, V * 3, "colors (V,3)") if colors is not None else None, V,
                 ptr(faces), T, ptr(keep), ptr(vref), ptr(voff), ptr(foff), ptr(verts_out), ptr(colors_out), ptr(faces_out), st)
        return verts_out, faces_out, colors_out

    def cull_by_frames(self, verts, faces, colors, c2ws, depths, eval_rec: bool):
        """cull_mesh: faces whose three vertices are all unseen go (cull_mesh.py:100-102)."""
        seen = self.seen_by_frames(verts, c2ws, depths, eval_rec)
        return self.filter_faces(verts, faces, colors, seen, require_all=False)

    def cull_by_hull(self, verts, faces, colors, planes):
        """cull_out_bound_mesh: faces with all three vertices inside the bound stay (cull_mesh.py:144-146)."""
        return self.filter_faces(verts, faces, colors, self.inside_hull(verts, planes), require_all=True)


def grid_axes(marching_cubes_bound, resolution: float, padding: float = 0.05):
    """The per-axis coordinates of Mesher.get_grid_uniform (Mesher.py:168-195): n = round((hi - lo + 2 * padding) / resolution)
    samples of np.linspace(lo - padding, hi + padding, n) in double, rounded to fp32 -- [x, y, z] CPU tensors.
    marching_cubes_bound: cfg['mapping']['marching_cubes_bound'] * scale, (3,2)."""
    b = np.asarray(marching_cubes_bound.cpu() if torch.is_tensor(marching_cubes_bound) else marching_cubes_bound, dtype=np.float64)
    axes = []
    for a in range(3):
        lo, hi = b[a, 0], b[a, 1]
        n = int(torch.tensor((hi - lo + 2 * padding) / resolution, dtype=torch.float64).round().int().item())
        axes.append(torch.from_numpy(np.linspace(lo - padding, hi + padding, n)).float())
    return axes


def get_mesh(mesh_out_file: Optional[str], meta: ops.FieldMeta, sdf_table, rgb_table, dec, bound, marching_cubes_bound, *,
             resolution: float = 0.01, level_set: float = 0.0, scale: float = 1.0, color: bool = True, mesh_bound=None,
             y_range: Optional[Tuple[int, int]] = None, keys: bool = False):
    """Mesher.get_mesh (Mesher.py:197-276) in one call on the device: get_grid_uniform -> eval_points over the whole grid
    (usl_sdf_query_grid) -> marching cubes (usl_mc_*) -> vertex colours (the fused field query) -> vertices / scale ->
    cull_out_bound_mesh against `mesh_bound` (usl_mesh_cull_hull + face rule + compaction) -> PLY.

    mesh_bound: what Mesher.get_bound_from_frames returns (an object with .vertices / .faces, i.e. the trimesh hull), or an
    (F,4) array of outward planes, or None (no bound culling: the reference's result before Mesher.py:274).
    y_range = (begin, end): only that y-slab of the grid (multi-GPU: parallel.slab_range; a slab that is not the last one
    queries one halo row); with keys=True the slab's vertex edge keys are returned as a fourth value for mesh.weld.
    Returns (verts (V,3) fp32, faces (T,3) int32, colors (V,3) uint8 | None) on the device -- or None when the level set
    does not cross the volume (the reference prints 'marching_cubes error' and returns) -- and writes mesh_out_file if given."""
    from .steps import DenseSdfQuery
    axes = grid_axes(marching_cubes_bound, resolution)
    dev = sdf_table.device
    q = DenseSdfQuery(meta, sdf_table.detach(), rgb_table.detach(), [d.detach() for d in dec], [a.to(dev) for a in axes])
    yb, ye = (0, q.ny) if y_range is None else (int(y_range[0]), int(y_range[1]))
    yh = min(ye + 1, q.ny)
    vol = q.run(yb, yh).view(yh - yb, q.nx, q.nz)
    out = MeshExtractor(axes, level_set).run(vol, yb, ye, halo=yh > ye, keys=True)
    verts, faces, vkeys = out
    if verts.shape[0] == 0 or faces.shape[0] == 0:
        return None
    cols = vertex_colors(meta, sdf_table, rgb_table, dec, verts, bound) if color else None     # on the un-scaled vertices (Mesher.py:259-267)
    if scale != 1.0:
        verts = verts / scale                                                                   # Mesher.py:271
    if mesh_bound is not None:
        planes = hull_planes(mesh_bound.vertices, mesh_bound.faces) if hasattr(mesh_bound, "vertices") else np.asarray(mesh_bound, dtype=np.float32)
        kept = []
        verts, faces, cols = MeshCuller.filter_faces(verts, faces, cols, MeshCuller.inside_hull(verts, planes), True, vertex_kept=kept)
        vkeys = vkeys[kept[0].bool()]
    if mesh_out_file is not None:
        write_ply(mesh_out_file, verts.cpu().numpy(), faces.cpu().numpy(), cols.cpu().numpy() if cols is not None else None)
    return (verts, faces, cols, vkeys) if keys else (verts, faces, cols)


def weld(parts):
    """Concatenate per-slab meshes [(verts, faces, keys, colors|None), ...] and weld the vertices that neighbouring slabs both
    emitted (the x / z edges of a halo row) by their global edge key.  Host numpy; returns verts, faces, colors."""
    vs = np.concatenate([p[0].cpu().numpy() for p in parts], axis=0)
    ks = np.concatenate([p[2].cpu().numpy() for p in parts], axis=0)
    cs = np.concatenate([p[3].cpu().numpy() for p in parts], axis=0) if parts[0][3] is not None else None
    offs = np.cumsum([0] + [p[0].shape[0] for p in parts[:-1]])
    fs = np.concatenate([p[1].cpu().numpy().astype(np.int64) + o for p, o in zip(parts, offs)], axis=0)
    uk, first, inv = np.unique(ks, return_index=True, return_inverse=True)
    return vs[first], inv[fs], (cs[first] if cs is not None else None)


def cull_by_bound(verts: np.ndarray, faces: np.ndarray, colors, lo, hi):
    """Keep the faces whose three vertices lie inside [lo, hi] and drop unreferenced vertices (the axis-aligned part of
    cull_out_bound_mesh, src/tools/cull_mesh.py:118-148; the reference's bound is a convex hull built with open3d)."""
    inside = np.all((verts >= np.asarray(lo)) & (verts <= np.asarray(hi)), axis=1)
    keep = inside[faces].all(axis=1)
    faces = faces[keep]
    used = np.unique(faces)
    remap = -np.ones(len(verts), dtype=np.int64); remap[used] = np.arange(len(used))
    return verts[used], remap[faces], (colors[used] if colors is not None else None)


def write_ply(path: str, verts: np.ndarray, faces: np.ndarray, colors: Optional[np.ndarray] = None, scale: float = 1.0):
    """Binary little-endian PLY as trimesh.Trimesh(vertices / scale, faces, vertex_colors).export writes it (Mesher.py:271-276)."""
    v = (np.asarray(verts, dtype=np.float64) / scale).astype("<f4")
    f = np.asarray(faces, dtype="<i4")
    hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {len(v)}", "property float x", "property float y", "property float z"]
    if colors is not None:
        hdr += ["property uchar red", "property uchar green", "property uchar blue", "property uchar alpha"]
    hdr += [f"element face {len(f)}", "property list uchar int vertex_indices", "end_header"]
    with open(path, "wb") as fh:
        fh.write(("\n".join(hdr) + "\n").encode("ascii"))
        if colors is not None:
            rec = np.zeros(len(v), dtype=[("p", "<f4", 3), ("c", "u1", 4)])
            rec["p"] = v; rec["c"][:, :3] = np.asarray(colors, dtype=np.uint8); rec["c"][:, 3] = 255
        else:
            rec = np.zeros(len(v), dtype=[("p", "<f4", 3)])
            rec["p"] = v
        fh.write(rec.tobytes())
        frec = np.zeros(len(f), dtype=[("n", "u1"), ("i", "<i4", 3)])
        frec["n"] = 3; frec["i"] = f
        fh.write(frec.tobytes())


def read_ply(path: str):
    """Reader for the files write_ply produces (tests / round trips)."""
    with open(path, "rb") as fh:
        nv = nf = 0
        has_c = False
        while True:
            line = fh.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                nv = int(line.split()[-1])
            elif line.startswith("element face"):
                nf = int(line.split()[-1])
            elif line == "property uchar red":
                has_c = True
            elif line == "end_header":
                break
        vd = [("p", "<f4", 3)] + ([("c", "u1", 4)] if has_c else [])
        v = np.frombuffer(fh.read(nv * np.dtype(vd).itemsize), dtype=vd)
        f = np.frombuffer(fh.read(nf * 13), dtype=[("n", "u1"), ("i", "<i4", 3)])
    return v["p"].copy(), f["i"].copy(), (v["c"][:, :3].copy() if has_c else None)
