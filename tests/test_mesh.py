"""f4: marching cubes on the device-resident SDF volume (csrc/mesh.cu, uni-slam_b200/mesh.py) against the oracle
(oracle/mc_ref.py) and against size-independent properties of a correct surface extraction."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import mc_ref
from helpers import pkg

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


def _sphere(n=(40, 36, 44), r=0.8):
    ax = [np.linspace(-1.2, 1.2, k).astype(np.float32) for k in n]            # x, y, z axes
    X, Y, Z = np.meshgrid(ax[0], ax[1], ax[2], indexing="xy")                  # shapes (ny, nx, nz): the dense query's layout
    return ax, (np.sqrt(X * X + Y * Y + Z * Z) - r).astype(np.float32)


# ---- CPU: the oracle itself and the generated header -----------------------------------------------------------------
def test_case_table_properties():
    t = mc_ref.tri_table()
    assert len(t) == 256 and not t[0] and not t[255] and mc_ref.max_triangles() == 5
    for cfg, tris in enumerate(t):
        crossing = {e for e in range(12) if ((cfg >> mc_ref.edge_corners(e)[0]) & 1) != ((cfg >> mc_ref.edge_corners(e)[1]) & 1)}
        assert {e for tri in tris for e in tri} == crossing, cfg              # every crossed edge carries a vertex, no other does
        de = [(a, b) for tri in tris for a, b in ((tri[0], tri[1]), (tri[1], tri[2]), (tri[2], tri[0]))]
        assert len(set(de)) == len(de), cfg                                    # oriented: no directed edge twice inside a cube


def test_committed_header_matches_the_construction():
    gen = importlib.machinery.SourceFileLoader("gen_mc_tables", os.path.join(REPO, "tools", "gen_mc_tables.py")).load_module()
    assert open(os.path.join(REPO, "uni-slam_b200", "csrc", "mc_tables.h")).read() == gen.render()


def test_oracle_sphere_is_a_closed_surface_on_the_level_set():
    ax, vol = _sphere()
    sp = [a[2] - a[1] for a in ax]
    v, f, _ = mc_ref.marching_cubes(vol, 0.0, [a[0] for a in ax], sp)
    closed, chi = mc_ref.mesh_stats(v, f)
    assert closed and chi == 2                                                 # watertight, oriented, genus 0
    assert np.abs(np.linalg.norm(v, axis=1) - 0.8).max() < 0.5 * max(sp) ** 2  # linear interpolation error is O(h^2)
    p0, p1, p2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    assert (np.einsum("ij,ij->i", np.cross(p1 - p0, p2 - p0), p0 + p1 + p2) > 0).all()   # normals towards larger values
    rng = np.random.default_rng(1)                                             # every ambiguous configuration: still crack-free
    vol = np.pad(rng.standard_normal((12, 11, 13)).astype(np.float32), 1, constant_values=5.0)
    v, f, _ = mc_ref.marching_cubes(vol, 0.0, (0, 0, 0), (1, 1, 1))
    assert mc_ref.mesh_stats(v, f)[0]


def test_ply_round_trip(tmp_path):
    mesh = importlib.import_module("uni-slam_b200.mesh")
    ax, vol = _sphere((12, 13, 11))
    v, f, _ = mc_ref.marching_cubes(vol, 0.0, [a[0] for a in ax], [a[2] - a[1] for a in ax])
    c = (np.abs(v) * 200).astype(np.uint8)
    p = str(tmp_path / "m.ply")
    mesh.write_ply(p, v, f, c, scale=2.0)
    v2, f2, c2 = mesh.read_ply(p)
    assert np.array_equal(v2, (v.astype(np.float64) / 2.0).astype(np.float32)) and np.array_equal(f2, f) and np.array_equal(c2, c)
    v3, f3, c3 = mesh.cull_by_bound(v, f, c, [-2, -2, 0.0], [2, 2, 2])        # upper half only
    assert len(f3) < len(f) and v3[:, 2].min() >= 0.0 and f3.max() == len(v3) - 1


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sphere", "random"])
def test_gpu_marching_cubes_matches_oracle_bit_for_bit(kind):
    mesh = importlib.import_module("uni-slam_b200.mesh")
    if kind == "sphere":
        ax, vol = _sphere()
    else:
        rng = np.random.default_rng(3)
        vol = rng.standard_normal((23, 19, 21)).astype(np.float32)            # (ny, nx, nz)
        ax = [np.linspace(0, 1, 19).astype(np.float32), np.linspace(0, 2, 23).astype(np.float32), np.linspace(-1, 0, 21).astype(np.float32)]
    ex = mesh.MeshExtractor([torch.from_numpy(a) for a in ax], level=0.0)
    v_ref, f_ref, k_ref = mc_ref.marching_cubes(vol, 0.0, ex.origin, ex.spacing)
    v, f, k = ex.run(torch.from_numpy(vol).to(DEV).contiguous(), keys=True)
    assert np.array_equal(v.cpu().numpy(), v_ref)                              # same vertices, same order, same bits
    assert np.array_equal(f.cpu().numpy().astype(np.int64), f_ref)
    assert np.array_equal(k.cpu().numpy(), k_ref)


@pytest.mark.gpu
def test_gpu_marching_cubes_slabs_weld_to_the_single_slab_mesh():
    """y-slabs with a halo row (the multi-GPU sharding of the dense query): per-slab meshes welded by edge key == the mesh of
    the whole volume (same vertex set, same triangles)."""
    mesh = importlib.import_module("uni-slam_b200.mesh")
    par = importlib.import_module("uni-slam_b200.parallel")
    ax, vol = _sphere((30, 37, 26))
    ex = mesh.MeshExtractor([torch.from_numpy(a) for a in ax])
    full = torch.from_numpy(vol).to(DEV).contiguous()
    v0, f0, k0 = ex.run(full, keys=True)
    ny, parts = vol.shape[0], []
    for r in range(3):
        yb, ye = par.slab_range(ny, r, 3)
        halo = ye < ny
        parts.append(ex.run(full[yb:ye + (1 if halo else 0)].contiguous(), yb, ye, halo=halo, keys=True) + (None,))
    v, f, _ = mesh.weld(parts)
    # canonical form: triangles as triples of edge keys
    keys_full = k0.cpu().numpy()
    kk = np.unique(np.concatenate([p[2].cpu().numpy() for p in parts]))
    assert np.array_equal(np.sort(keys_full), kk)
    tri_full = np.sort(keys_full[f0.cpu().numpy()].reshape(-1, 3), axis=0)
    tri_parts = np.sort(kk[f].reshape(-1, 3), axis=0)
    assert tri_full.shape == tri_parts.shape
    canon = lambda t: np.unique(t.view([("a", t.dtype), ("b", t.dtype), ("c", t.dtype)]))
    assert np.array_equal(canon(np.ascontiguousarray(keys_full[f0.cpu().numpy()])), canon(np.ascontiguousarray(kk[f])))
    closed, chi = mc_ref.mesh_stats(v, f)
    assert closed and chi == 2


@pytest.mark.gpu
def test_mesh_from_the_field_end_to_end(tmp_path):
    """Dense SDF query -> marching cubes -> vertex colours -> PLY on a small grid of the golden field: every vertex sits on a
    sign change of the queried volume and on the field's zero level (|sdf(vertex)| small), colours come from the colour field."""
    P = pkg()
    mesh = importlib.import_module("uni-slam_b200.mesh")
    import gpu_cases
    from helpers import load_golden
    g = load_golden("mesh_replica")
    meta, tabs, dec, beta = gpu_cases.cuda_field(g, 120, DEV)
    lo, hi = g["bound"][:, 0], g["bound"][:, 1]
    axes = [torch.linspace(float(lo[a]) + 0.3, float(hi[a]) - 0.3, n) for a, n in zip(range(3), (41, 33, 27))]
    q = P.DenseSdfQuery(meta, tabs[0], tabs[1], dec, [a.to(DEV) for a in axes])
    vol = q.run(0, q.ny).reshape(q.ny, q.nx, q.nz)
    ex = mesh.MeshExtractor(axes)
    v, f = ex.run(vol.contiguous())
    assert v.shape[0] > 100 and f.shape[0] > 100
    v_ref, f_ref, _ = mc_ref.marching_cubes(vol.cpu().numpy(), 0.0, ex.origin, ex.spacing)
    assert np.array_equal(v.cpu().numpy(), v_ref) and np.array_equal(f.cpu().numpy().astype(np.int64), f_ref)
    col = mesh.vertex_colors(meta, tabs[0], tabs[1], dec, v, torch.from_numpy(g["bound"]))
    assert col.shape == (v.shape[0], 3) and col.dtype == torch.uint8 and int(col.max()) > int(col.min())
    p = str(tmp_path / "scene.ply")
    mesh.write_ply(p, v.cpu().numpy(), f.cpu().numpy(), col.cpu().numpy())
    v2, f2, c2 = mesh.read_ply(p)
    assert np.array_equal(v2, v.cpu().numpy()) and np.array_equal(c2, col.cpu().numpy())
