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
not in this image).  The frustum / bound culling the reference applies afterwards (cull_mesh.py) needs the dataset frames
and open3d and stays host policy; `cull_by_bound` below is the bound part of it on the device arrays.
"""
from ctypes import byref, c_int64
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import ops
from ._lib import call, ptr, stream


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
    """Binary little-endian PLY as trimesh.Trimesh(vertices / scale, faces, vertex_colors).export writes it (Mesher.py:269-276)."""
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
