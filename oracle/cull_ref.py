"""oracle/cull_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the mesh culling the reference applies after marching cubes:

* ``visibility``      src/tools/cull_mesh.py:58-99 (cull_mesh): per frame, project every vertex with torch.inverse(c2w) and the
                      pinhole matrix (x flipped, z + 1e-5), sample the sensor depth bilinearly (F.grid_sample, align_corners,
                      zero padding, on the reference's u/W, v/H normalisation), and mark the vertex seen when it lies in the
                      frustum in front of the camera (and, with eval_rec, not further than depth + truncation).
* ``face_filter``     cull_mesh.py:100-103 / :144-146: the faces trimesh's update_faces keeps and the vertex compaction of
                      remove_unreferenced_vertices (order-preserving).
* ``hull_planes`` / ``inside_hull``  the point-in-convex-hull test behind ``mesh_bound.contains`` (cull_mesh.py:137-143).
                      trimesh is absent from this image: for a closed convex hull its ray test equals the half-space test
                      up to points ON the hull (parity unpinned for those).

Pinned by tests/golden/cull_replica.npz, which oracle/gen_golden.py writes by running the UNMODIFIED cull_mesh /
cull_out_bound_mesh on a small mesh through stand-ins for trimesh's container methods and the dataset reader.

``visibility`` also returns, per vertex, the smallest distance of any of its comparisons to its decision threshold
(in the comparison's own unit: pixels or metres), so that a test can demand exact agreement wherever the decision is
not within rounding of a tie.
"""
import numpy as np
import torch
import torch.nn.functional as F


def project(verts32: torch.Tensor, c2w: torch.Tensor, fx, fy, cx, cy):
    """cull_mesh.py:69-84.  verts32 (V,3) fp32, c2w (4,4) fp32 -> u, v, z (V,) fp32 with z = K-row-3 . cam + 1e-5."""
    w2c = torch.inverse(c2w)
    hom = torch.cat([verts32, torch.ones_like(verts32[:, :1])], dim=1)
    cam = (w2c @ hom[:, :, None])[:, :3, 0]
    cam = torch.stack([-cam[:, 0], cam[:, 1], cam[:, 2]], dim=1)
    Kmat = torch.tensor([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=torch.float64).float()
    uvw = (Kmat @ cam[:, :, None])[:, :, 0]
    z = uvw[:, 2] + 1e-5
    return uvw[:, 0] / z, uvw[:, 1] / z, z


def sample_depth(depth: torch.Tensor, u: torch.Tensor, v: torch.Tensor):
    """cull_mesh.py:86-90: the sensor depth at (u/W, v/H) mapped to [-1,1], bilinear, align_corners, zero padding."""
    H, W = depth.shape
    g = torch.stack([u / W, v / H], dim=-1)[None, None]
    g = 2 * g - 1
    return F.grid_sample(depth[None, None], g, padding_mode="zeros", align_corners=True).reshape(-1)


def visibility(verts, c2ws, depths, fx, fy, cx, cy, truncation, eval_rec, edge=0):
    """seen (V,) bool = NOT whole_mask of cull_mesh.py:58-99, margin (V,) float64.
    verts (V,3) float (the PLY's vertices); c2ws (K,4,4) fp32; depths (K,H,W) fp32."""
    v32 = torch.as_tensor(np.asarray(verts)).float()
    V = v32.shape[0]
    seen = torch.zeros(V, dtype=torch.bool)
    margin = torch.full((V,), float("inf"), dtype=torch.float64)
    K, H, W = depths.shape
    for k in range(K):
        u, v, z = project(v32, c2ws[k].float(), fx, fy, cx, cy)
        ds = sample_depth(depths[k].float(), u, v)
        nz = -z
        terms = [(nz, 0.0, ">="), (u, float(W - edge), "<"), (u, float(edge), ">"), (v, float(H - edge), "<"), (v, float(edge), ">")]
        m = (0 <= nz) & (u < W - edge) & (u > edge) & (v < H - edge) & (v > edge)
        if eval_rec:
            m = m & (ds + truncation >= nz)
            terms.append((ds + truncation - nz, 0.0, ">="))
        seen |= m
        d = torch.stack([(a.double() - b).abs() for a, b, _ in terms], dim=0)
        d = torch.where(torch.isfinite(d), d, torch.full_like(d, float("inf")))
        margin = torch.minimum(margin, d.min(dim=0)[0])
    return seen.numpy(), margin.numpy()


def face_filter(verts, faces, vmask, require_all, colors=None):
    """update_faces(keep) + remove_unreferenced_vertices (cull_mesh.py:100-103, 144-146).
    require_all=False: keep a face when ANY of its vertices has vmask (cull_mesh: drop faces whose three vertices are all
    unseen); require_all=True: keep when ALL have it (cull_out_bound_mesh).  Order-preserving.  Returns verts, faces, colors, keep."""
    faces = np.asarray(faces)
    m = np.asarray(vmask, dtype=bool)[faces]
    keep = m.all(axis=1) if require_all else m.any(axis=1)
    f = faces[keep]
    ref = np.zeros(len(verts), dtype=bool)
    ref[f.reshape(-1)] = True
    remap = np.cumsum(ref) - 1
    return np.asarray(verts)[ref], remap[f], (np.asarray(colors)[ref] if colors is not None else None), keep


def hull_planes(hull_verts, hull_faces):
    """(F,4) float64 outward planes n.x + d <= 0 inside, unit normals, oriented by the hull's vertex centroid."""
    hv = np.asarray(hull_verts, dtype=np.float64); hf = np.asarray(hull_faces)
    a, b, c = hv[hf[:, 0]], hv[hf[:, 1]], hv[hf[:, 2]]
    n = np.cross(b - a, c - a)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    d = -(n * a).sum(axis=1)
    flip = (n @ hv.mean(axis=0) + d) > 0
    n[flip] *= -1; d[flip] *= -1
    return np.concatenate([n, d[:, None]], axis=1)


def inside_hull(points, planes):
    """inside (V,) bool and margin (V,) = distance of the point to the nearest hull plane."""
    s = np.asarray(points, dtype=np.float64) @ planes[:, :3].T + planes[:, 3]
    return (s <= 0).all(axis=1), np.abs(s).min(axis=1)
