"""f4, the culling after marching cubes (src/tools/cull_mesh.py): csrc/cull.cu + mesh.MeshCuller.

Chain of evidence:
  unmodified cull_mesh / cull_out_bound_mesh (run by oracle/gen_golden.py) -> tests/golden/cull_replica.npz
  oracle/cull_ref.py == golden                                                    (CPU)
  thread functions of the kernels (csrc/usl_cull.cuh, compiled for the host by tests/host_harness) == golden, and == oracle on
      200 k random points wherever the decision is not within rounding of a tie  (CPU)
  CUDA kernels == host harness bit for bit, == golden, end-to-end culled meshes   (-m gpu)
"""
import numpy as np
import pytest
import torch

import helpers
from helpers import pkg
from oracle import cull_ref

DEV = "cuda:0"


def _golden():
    g = helpers.load_golden("cull_replica")
    H, W, fx, fy, cx, cy = [float(v) for v in g["meta_H_W_fx_fy_cx_cy"]]
    return g, (int(H), int(W), fx, fy, cx, cy), float(g["truncation"])


def _random_problem(n_pts=200000, K=21, seed=5):
    """Points in and around the golden's room, the golden's frames repeated with jittered poses."""
    g, cam, tr = _golden()
    rng = np.random.default_rng(seed)
    lo, hi = g["verts"].min(axis=0) - 0.5, g["verts"].max(axis=0) + 0.5
    pts = (lo + (hi - lo) * rng.random((n_pts, 3))).astype(np.float32)
    c2ws, depths = [], []
    for k in range(K):
        c = g["c2ws"][k % len(g["c2ws"])].copy()
        c[:3, 3] += rng.normal(0, 0.05, 3).astype(np.float32)
        c2ws.append(c); depths.append(g["depths"][k % len(g["depths"])])
    return pts, np.stack(c2ws).astype(np.float32), np.stack(depths).astype(np.float32), cam, tr


# ---- CPU -------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,eval_rec", [("rec", True), ("vis", False)])
def test_oracle_matches_the_unmodified_reference(tag, eval_rec):
    g, cam, tr = _golden()
    seen, margin = cull_ref.visibility(g["verts"], torch.from_numpy(g["c2ws"]), torch.from_numpy(g["depths"]), *cam[2:], tr, eval_rec)
    assert np.array_equal(seen, g["seen_" + tag])
    v, f, _, keep = cull_ref.face_filter(g["verts"], g["faces"], seen, require_all=False)
    assert np.array_equal(keep, g["face_keep_" + tag]) and np.array_equal(v, g["culled_verts_" + tag]) and np.array_equal(f, g["culled_faces_" + tag])
    assert 0 < seen.sum() < len(seen) and 0 < keep.sum() < len(keep)          # the case exercises both outcomes


def test_oracle_hull_matches_the_unmodified_reference():
    g, _, _ = _golden()
    planes = cull_ref.hull_planes(g["hull_verts"], g["hull_faces"])
    inside, _ = cull_ref.inside_hull(g["verts"], planes)
    assert np.array_equal(inside, g["inside_hull"]) and 0 < inside.sum() < len(inside)
    v, f, _, keep = cull_ref.face_filter(g["verts"], g["faces"], inside, require_all=True)
    assert np.array_equal(keep, g["face_keep_hull"]) and np.array_equal(v, g["culled_verts_hull"]) and np.array_equal(f, g["culled_faces_hull"])
    # the product's host-side plane construction (fp32) is the oracle's (fp64) rounded
    pl = pkg().mesh.hull_planes(g["hull_verts"], g["hull_faces"])
    assert pl.dtype == np.float32 and np.allclose(pl, planes, atol=1e-6)


@pytest.mark.parametrize("tag,eval_rec", [("rec", True), ("vis", False)])
def test_kernel_element_functions_reproduce_the_reference_masks(tag, eval_rec):
    """usl_cull.cuh compiled for the host: the same masks and the same culled mesh as the unmodified reference, for every
    grouping of the frames (the kernel's early exit / OR accumulation must not change the result)."""
    g, cam, tr = _golden()
    w2c = torch.inverse(torch.from_numpy(g["c2ws"])).numpy()
    for fpc, nthreads in ((16, 256), (3, 768), (1, 1024), (64, 4736 * 256)):     # grid-stride loops of every length, one group .. one per frame
        seen = helpers.cull_host_frames(g["verts"], w2c, g["depths"], cam, tr, eval_rec, fpc, nthreads)
        assert np.array_equal(seen.astype(bool), g["seen_" + tag]), (fpc, nthreads)
    col = (np.arange(len(g["verts"]) * 3) % 251).astype(np.uint8).reshape(-1, 3)
    for nthreads in (256, 2304, 4736 * 256):
        v, f, c, keep = helpers.cull_host_compact(g["verts"], col, g["faces"], seen, require_all=0, nthreads=nthreads)
        assert np.array_equal(keep.astype(bool), g["face_keep_" + tag])
        assert np.array_equal(v, g["culled_verts_" + tag]) and np.array_equal(f, g["culled_faces_" + tag])
    _, _, c_ref, _ = cull_ref.face_filter(g["verts"], g["faces"], seen.astype(bool), False, colors=col)
    assert np.array_equal(c, c_ref)
    planes = pkg().mesh.hull_planes(g["hull_verts"], g["hull_faces"])
    for nthreads in (256, 1280, 4736 * 256):
        inside = helpers.cull_host_hull(g["verts"], planes, nthreads)
        assert np.array_equal(inside.astype(bool), g["inside_hull"])
    v, f, _, _ = helpers.cull_host_compact(g["verts"], None, g["faces"], inside, require_all=1)
    assert np.array_equal(v, g["culled_verts_hull"]) and np.array_equal(f, g["culled_faces_hull"])


@pytest.mark.parametrize("eval_rec", [True, False])
def test_kernel_element_functions_match_the_oracle_on_random_points(eval_rec):
    pts, c2ws, depths, cam, tr = _random_problem()
    seen_ref, margin = cull_ref.visibility(pts, torch.from_numpy(c2ws), torch.from_numpy(depths), *cam[2:], tr, eval_rec)
    seen = helpers.cull_host_frames(pts, torch.inverse(torch.from_numpy(c2ws)).numpy(), depths, cam, tr, eval_rec, 16).astype(bool)
    diff = seen != seen_ref
    # a disagreement is only admissible where some comparison sits within rounding of its threshold (pixels / metres)
    assert not (diff & (margin > 1e-4)).any(), (int(diff.sum()), float(margin[diff].max()) if diff.any() else 0.0)
    assert diff.sum() <= 1e-4 * len(pts)
    assert 0.05 < seen.mean() < 0.95


def test_face_rule_edge_cases():
    v = np.zeros((4, 3), dtype=np.float32); v[:, 0] = np.arange(4)
    f = np.array([[0, 1, 2], [1, 2, 3], [0, 0, 0], [3, 3, 3], [0, 1, 9], [-1, 1, 2]], dtype=np.int32)
    m = np.array([1, 0, 0, 0], dtype=np.uint8)
    vo, fo, _, keep = helpers.cull_host_compact(v, None, f, m, require_all=0)
    assert keep.tolist() == [1, 0, 1, 0, 0, 0]                                   # out-of-range indices drop the face
    assert np.array_equal(vo, v[:3]) and fo.tolist() == [[0, 1, 2], [0, 0, 0]]
    vo, fo, _, keep = helpers.cull_host_compact(v, None, f, m, require_all=1)
    assert keep.tolist() == [0, 0, 1, 0, 0, 0] and len(vo) == 1 and fo.tolist() == [[0, 0, 0]]
    vo, fo, _, keep = helpers.cull_host_compact(v, None, f[:0], m, require_all=0)
    assert len(vo) == 0 and len(fo) == 0


# ---- GPU -------------------------------------------------------------------------------------------------------------
def _gpu_seen(culler, pts, c2ws, depths, eval_rec):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    seen = culler.seen_by_frames(t(pts), t(c2ws), t(depths), eval_rec)
    torch.cuda.synchronize()
    return seen.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("eval_rec", [True, False])
def test_gpu_cull_frames_matches_harness_and_golden(eval_rec):
    P = pkg()
    g, cam, tr = _golden()
    culler = P.mesh.MeshCuller(*cam, tr)
    w2c = torch.inverse(torch.from_numpy(g["c2ws"]).to(DEV)).cpu().numpy()        # the matrices the kernel is given
    ref = helpers.cull_host_frames(g["verts"], w2c, g["depths"], cam, tr, eval_rec, 16)
    for fpc in (0, 3, 1, 64):
        culler.frames_per_cta = fpc
        seen = _gpu_seen(culler, g["verts"], g["c2ws"], g["depths"], eval_rec)
        assert np.array_equal(seen, ref), fpc                                      # bit for bit: same source, same rounding
    _, margin = cull_ref.visibility(g["verts"], torch.from_numpy(g["c2ws"]), torch.from_numpy(g["depths"]), *cam[2:], tr, eval_rec)
    diff = seen.astype(bool) != g["seen_rec" if eval_rec else "seen_vis"]
    assert not (diff & (margin > 1e-4)).any()                                      # the unmodified reference's mask


@pytest.mark.gpu
@pytest.mark.parametrize("eval_rec", [True, False])
def test_gpu_cull_frames_random_points(eval_rec):
    P = pkg()
    pts, c2ws, depths, cam, tr = _random_problem(n_pts=300001, K=37)
    culler = P.mesh.MeshCuller(*cam, tr)
    w2c = torch.inverse(torch.from_numpy(c2ws).to(DEV)).cpu().numpy()
    ref = helpers.cull_host_frames(pts, w2c, depths, cam, tr, eval_rec, 16)
    seen = _gpu_seen(culler, pts, c2ws, depths, eval_rec)
    assert np.array_equal(seen, ref)
    seen_ref, margin = cull_ref.visibility(pts, torch.from_numpy(c2ws), torch.from_numpy(depths), *cam[2:], tr, eval_rec)
    diff = seen.astype(bool) != seen_ref
    assert not (diff & (margin > 1e-4)).any() and diff.sum() <= 1e-4 * len(pts)
    # accumulation over ranges of frames == one call (the two calls invert their poses in batches of another size, which may
    # round a matrix differently: agreement is required wherever no comparison is within 1e-4 of a tie)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    acc = culler.seen_by_frames(t(pts), t(c2ws[:20]), t(depths[:20]), eval_rec)
    acc = culler.seen_by_frames(t(pts), t(c2ws[20:]), t(depths[20:]), eval_rec, seen=acc)
    diff = acc.cpu().numpy() != ref
    assert not (diff & (margin > 1e-4)).any() and diff.sum() <= 1e-4 * len(pts)


@pytest.mark.gpu
def test_gpu_culled_meshes_match_the_reference():
    P = pkg()
    g, cam, tr = _golden()
    culler = P.mesh.MeshCuller(*cam, tr)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    verts, faces = t(g["verts"]), t(g["faces"].astype(np.int32))
    col = (np.arange(len(g["verts"]) * 3) % 251).astype(np.uint8).reshape(-1, 3)
    for tag, eval_rec in (("rec", True), ("vis", False)):
        v, f, c = culler.cull_by_frames(verts, faces, t(col), t(g["c2ws"]), t(g["depths"]), eval_rec)
        torch.cuda.synchronize()
        seen = culler.seen_by_frames(verts, t(g["c2ws"]), t(g["depths"]), eval_rec).cpu().numpy().astype(bool)
        v_ref, f_ref, c_ref, _ = cull_ref.face_filter(g["verts"], g["faces"], seen, False, colors=col)
        assert np.array_equal(v.cpu().numpy(), v_ref) and np.array_equal(f.cpu().numpy(), f_ref) and np.array_equal(c.cpu().numpy(), c_ref)
        if np.array_equal(seen, g["seen_" + tag]):                                  # (it is, unless a tie rounds differently on the device)
            assert np.array_equal(v.cpu().numpy(), g["culled_verts_" + tag]) and np.array_equal(f.cpu().numpy(), g["culled_faces_" + tag])
    planes = P.mesh.hull_planes(g["hull_verts"], g["hull_faces"])
    inside = culler.inside_hull(verts, planes).cpu().numpy()
    assert np.array_equal(inside, helpers.cull_host_hull(g["verts"], planes))
    assert np.array_equal(inside.astype(bool), g["inside_hull"])
    v, f, c = culler.cull_by_hull(verts, faces, None, planes)
    assert c is None and np.array_equal(v.cpu().numpy(), g["culled_verts_hull"]) and np.array_equal(f.cpu().numpy(), g["culled_faces_hull"])
    # a hull with more planes than one shared-memory tile, and an empty selection
    big = np.concatenate([planes] * 5, axis=0)
    assert np.array_equal(culler.inside_hull(verts, big).cpu().numpy(), inside)
    far = culler.cull_by_hull(verts + 100.0, faces, None, planes)
    assert far[0].shape[0] == 0 and far[1].shape[0] == 0


# ---- CPU: the glue of mesh.get_mesh (Mesher.get_mesh in one call) with host stand-ins for the launches --------------------
def test_grid_axes_match_the_reference():
    """mesh.grid_axes against the axes the unmodified Mesher.get_grid_uniform produced (tests/golden/mesh_*.npz)."""
    for name in ("mesh_replica", "mesh_scannet"):
        g = helpers.load_golden(name)
        axes = pkg().mesh.grid_axes(g["mc_bound"], float(g["resolution"]))
        for a, k in zip(axes, ("axis_x", "axis_y", "axis_z")):
            assert a.dtype == torch.float32 and np.array_equal(a.numpy(), g[k].astype(np.float32))


def test_get_mesh_glue_with_host_stand_ins(monkeypatch, tmp_path):
    """get_mesh's order of operations (query -> marching cubes -> colours on un-scaled vertices -> / scale -> bound culling ->
    PLY) with every launch replaced by its host counterpart of the SAME signature (oracle marching cubes, the culling
    harness); the launches themselves are covered by the -m gpu tests."""
    import importlib
    import inspect
    import types
    from oracle import mc_ref
    mesh = pkg().mesh
    steps = importlib.import_module("uni-slam_b200.steps")

    class Query:
        def __init__(self, meta, sdf_table, rgb_table, dec, axes):
            self.ax, self.ay, self.az = [a.double() for a in axes]
            self.nx, self.ny, self.nz = [a.numel() for a in axes]

        def run(self, y_begin, y_end, out=None):
            gy, gx, gz = torch.meshgrid(self.ay[y_begin:y_end], self.ax, self.az, indexing="ij")
            return (torch.sqrt((gx - 0.2) ** 2 + gy ** 2 + (gz + 0.1) ** 2) - 0.9).float().reshape(-1)   # a sphere: sdf < 0 inside

    assert list(inspect.signature(Query.__init__).parameters) == list(inspect.signature(steps.DenseSdfQuery.__init__).parameters)
    assert list(inspect.signature(Query.run).parameters) == list(inspect.signature(steps.DenseSdfQuery.run).parameters)

    def mc_run(self, vol, y_begin=0, y_end=None, halo=False, keys=False):
        assert vol.shape == (self.ny, self.nx, self.nz) and y_begin == 0 and not halo
        v, f, k = mc_ref.marching_cubes(vol.numpy(), self.level, self.origin, self.spacing)
        out = (torch.from_numpy(v), torch.from_numpy(f.astype(np.int32)), torch.from_numpy(k))
        return out if keys else out[:2]

    def colors(meta, sdf_table, rgb_table, dec, verts, bound):
        return (verts * 40).abs().clamp(0, 255).to(torch.uint8)                 # depends on the UN-scaled position

    def inside_hull(verts, planes):
        return torch.from_numpy(helpers.cull_host_hull(verts.numpy(), planes))

    def filter_faces(verts, faces, colors, vmask, require_all, vertex_kept=None):
        v, f, c, _ = helpers.cull_host_compact(verts.numpy(), colors.numpy() if colors is not None else None, faces.numpy(), vmask.numpy(), require_all)
        if vertex_kept is not None:
            ref = np.zeros(len(verts), dtype=np.uint8); ref[np.unique(faces.numpy()[cull_ref.face_filter(verts.numpy(), faces.numpy(), vmask.numpy(), require_all)[3]])] = 1
            vertex_kept.append(torch.from_numpy(ref))
        return torch.from_numpy(v), torch.from_numpy(f), (torch.from_numpy(c) if c is not None else None)

    for fake, real in ((mc_run, mesh.MeshExtractor.run), (colors, mesh.vertex_colors), (inside_hull, mesh.MeshCuller.inside_hull),
                       (filter_faces, mesh.MeshCuller.filter_faces)):
        assert list(inspect.signature(fake).parameters) == list(inspect.signature(real).parameters), real
    monkeypatch.setattr(steps, "DenseSdfQuery", Query)
    monkeypatch.setattr(mesh.MeshExtractor, "run", mc_run)
    monkeypatch.setattr(mesh, "vertex_colors", colors)
    monkeypatch.setattr(mesh.MeshCuller, "inside_hull", staticmethod(inside_hull))
    monkeypatch.setattr(mesh.MeshCuller, "filter_faces", staticmethod(filter_faces))

    mc_bound = np.array([[-1.2, 1.2], [-1.1, 1.3], [-1.4, 1.0]])
    tab = torch.zeros(4)
    scale = 2.0
    # the bound: a box that cuts the (scaled) sphere, as 12 triangles
    lo, hi = np.array([-0.6, -0.3, -0.6]), np.array([0.35, 0.6, 0.3])
    hv = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    hf = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]])
    ply = str(tmp_path / "m.ply")
    v, f, c, k = mesh.get_mesh(ply, None, tab, tab, [], None, mc_bound, resolution=0.1, scale=scale, mesh_bound=types.SimpleNamespace(vertices=hv, faces=hf), keys=True)
    # expected, step by step
    axes = mesh.grid_axes(mc_bound, 0.1)
    vol = Query(None, None, None, None, axes).run(0, axes[1].numel()).reshape(axes[1].numel(), axes[0].numel(), axes[2].numel())
    ev, ef, ek = mc_ref.marching_cubes(vol.numpy(), 0.0, [float(a[0]) for a in axes], [float(a[2] - a[1]) for a in axes])
    ec = colors(None, None, None, None, torch.from_numpy(ev), None).numpy()
    evs = (torch.from_numpy(ev) / scale).numpy()
    inside, _ = cull_ref.inside_hull(evs, cull_ref.hull_planes(hv, hf))
    rv, rf, rc, keep = cull_ref.face_filter(evs, ef, inside, True, colors=ec)
    assert 0 < len(rv) < len(ev) and 0 < len(rf) < len(ef)
    assert np.array_equal(v.numpy(), rv) and np.array_equal(f.numpy(), rf) and np.array_equal(c.numpy(), rc)
    ref = np.zeros(len(ev), dtype=bool); ref[np.unique(ef[keep])] = True
    assert np.array_equal(k.numpy(), ek[ref])
    v2, f2, c2 = mesh.read_ply(ply)
    assert np.array_equal(v2, rv) and np.array_equal(f2, rf) and np.array_equal(c2, rc)
    # no bound, no colours, no file; and a level set that misses the volume
    v, f, c = mesh.get_mesh(None, None, tab, tab, [], None, mc_bound, resolution=0.1, color=False)
    assert c is None and np.array_equal(v.numpy(), ev) and np.array_equal(f.numpy(), ef)
    assert mesh.get_mesh(None, None, tab, tab, [], None, mc_bound, resolution=0.1, level_set=50.0) is None


@pytest.mark.gpu
def test_gpu_get_mesh_one_call(tmp_path):
    """mesh.get_mesh (Mesher.get_mesh in one call) == its pieces run one by one, with and without the bound culling."""
    import gpu_cases
    P = pkg()
    mesh = P.mesh
    g = helpers.load_golden("mesh_replica")
    meta, tabs, dec, beta = gpu_cases.cuda_field(g, 120, DEV)
    bound = torch.from_numpy(g["bound"])
    mc_bound = np.stack([g["bound"][:, 0] + 0.3, g["bound"][:, 1] - 0.3], axis=1)
    ply = str(tmp_path / "a.ply")
    res = mesh.get_mesh(ply, meta, tabs[0], tabs[1], dec, bound, mc_bound, resolution=0.2)
    assert res is not None
    v, f, c = res
    axes = mesh.grid_axes(mc_bound, 0.2)
    q = P.DenseSdfQuery(meta, tabs[0], tabs[1], dec, [a.to(DEV) for a in axes])
    vol = q.run(0, q.ny).view(q.ny, q.nx, q.nz)
    v0, f0 = mesh.MeshExtractor(axes).run(vol.contiguous())
    c0 = mesh.vertex_colors(meta, tabs[0], tabs[1], dec, v0, bound)
    assert v0.shape[0] > 100 and torch.equal(v, v0) and torch.equal(f, f0) and torch.equal(c, c0)
    v2, f2, c2 = mesh.read_ply(ply)
    assert np.array_equal(v2, v0.cpu().numpy()) and np.array_equal(f2, f0.cpu().numpy()) and np.array_equal(c2, c0.cpu().numpy())
    # bound culling: six planes of a box around the middle of the mesh
    lo = (v0.min(dim=0)[0] * 0.7 + v0.max(dim=0)[0] * 0.3).cpu().numpy(); hi = (v0.min(dim=0)[0] * 0.3 + v0.max(dim=0)[0] * 0.7).cpu().numpy()
    planes = np.array([[1, 0, 0, -hi[0]], [-1, 0, 0, lo[0]], [0, 1, 0, -hi[1]], [0, -1, 0, lo[1]], [0, 0, 1, -hi[2]], [0, 0, -1, lo[2]]], dtype=np.float32)
    v, f, c, k = mesh.get_mesh(None, meta, tabs[0], tabs[1], dec, bound, mc_bound, resolution=0.2, mesh_bound=planes, keys=True)
    inside = helpers.cull_host_hull(v0.cpu().numpy(), planes)
    rv, rf, rc, keep = cull_ref.face_filter(v0.cpu().numpy(), f0.cpu().numpy(), inside, True, colors=c0.cpu().numpy())
    assert 0 < len(rv) < v0.shape[0]
    assert np.array_equal(v.cpu().numpy(), rv) and np.array_equal(f.cpu().numpy(), rf) and np.array_equal(c.cpu().numpy(), rc)
    assert k.shape[0] == len(rv) and len(torch.unique(k)) == len(rv)


# ---- f3: eval_rendering's per-frame metrics (usl_render_metrics, steps.RenderMetrics) -------------------------------------------
def test_render_metrics_oracle_and_element_function_match_the_reference():
    """tests/golden/evalr_replica.npz comes from the unmodified eval_rendering (src/tools/eval_recon.py:235-307): the oracle,
    the kernel's thread function (host build) and RenderMetrics.finalize reproduce its per-frame mse and its result line."""
    from oracle import path_ref
    P = pkg()
    g = helpers.load_golden("evalr_replica")
    accs = []
    for i in range(len(g["frames"])):
        t = [torch.from_numpy(g[k][i]) for k in ("gt_color", "gt_depth", "color", "depth")]
        mse, psnr, l1 = path_ref.render_metrics(*t)
        assert float(mse) == g["mse"][i] and float(l1) == g["depth_l1"][i]                    # the reference's own float64 values
        acc = helpers.metrics_host(g["gt_color"][i], g["gt_depth"][i], g["color"][i], g["depth"][i])
        assert acc[2] == int((g["gt_depth"][i] > 0).sum())
        assert abs(acc[0] / (3 * acc[2]) / g["mse"][i] - 1) < 1e-12 and abs(acc[1] / acc[2] / g["depth_l1"][i] - 1) < 1e-12
        accs.append(acc)
    res = P.RenderMetrics.finalize(torch.from_numpy(np.stack(accs)))
    assert float(f"{res['avg_psnr']:.4f}") == float(g["avg_psnr"]) and float(f"{res['depth_l1_render']:.4f}") == float(g["depth_l1_render"])
    assert res["frames"] == 3 and torch.isnan(P.RenderMetrics.finalize(torch.zeros(1, 3, dtype=torch.float64))["psnr"]).all()


@pytest.mark.gpu
def test_gpu_render_metrics():
    P = pkg()
    g = helpers.load_golden("evalr_replica")
    rm = P.RenderMetrics(DEV, max_frames=8)
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    for i in range(len(g["frames"])):
        # as eval_rendering receives them: dataset colour float64 (H,W,3), render_img's depth float64 (H,W), colour fp32
        rm.add(t(g["color"][i]), t(g["depth"][i]), t(g["gt_color"][i]), t(g["gt_depth"][i]))
    torch.cuda.synchronize()
    acc = rm.acc[:rm.n].cpu().numpy()
    for i in range(len(g["frames"])):
        ref = helpers.metrics_host(g["gt_color"][i], g["gt_depth"][i], g["color"][i], g["depth"][i])
        assert acc[i, 2] == ref[2] and np.allclose(acc[i, :2], ref[:2], rtol=1e-12, atol=0)   # double sums, another order
    res = rm.result()
    assert np.allclose(res["mse"].numpy(), g["mse"], rtol=1e-12)
    assert float(f"{res['avg_psnr']:.4f}") == float(g["avg_psnr"]) and float(f"{res['depth_l1_render']:.4f}") == float(g["depth_l1_render"])
    # a full-size frame (more pixels than one grid pass), partly without depth
    rng = np.random.default_rng(3)
    n = 680 * 1200
    gc, c = rng.random((n, 3), dtype=np.float32), rng.random((n, 3), dtype=np.float32)
    gd = (rng.random(n, dtype=np.float32) * 4).astype(np.float32); gd[rng.random(n) < 0.1] = 0
    d = (gd + rng.normal(0, 0.05, n)).astype(np.float32)
    rm2 = P.RenderMetrics(DEV, max_frames=1)
    rm2.add(t(c), t(d), t(gc), t(gd))
    ref = helpers.metrics_host(gc, gd, c, d)
    got = rm2.acc[0].cpu().numpy()
    assert got[2] == ref[2] and np.allclose(got[:2], ref[:2], rtol=1e-11, atol=0)
    with pytest.raises(RuntimeError):
        rm2.add(t(c), t(d), t(gc), t(gd))


def test_cull_thread_functions_on_degenerate_vertices_and_empty_inputs():
    """Vertices where the projection degenerates (camera centre, z = 0 plane, NaN / inf coordinates, far away) and empty
    meshes / frame lists: the kernels' thread functions decide like the oracle (torch semantics: comparisons with NaN are false)."""
    g, cam, tr = _golden()
    c2w = g["c2ws"][:2]
    ctr = c2w[0][:3, 3]
    fwd = -c2w[0][:3, 2]                                   # the camera looks along -z of its frame
    special = np.stack([ctr, ctr + 1e-5 * fwd, ctr - 1e-5 * fwd, ctr + 1.0 * fwd, ctr - 1.0 * fwd, ctr + 1e6 * fwd,
                        np.array([np.nan, 0, 0]), np.array([np.inf, 0, 0]), np.array([0, -np.inf, 0]), np.zeros(3)]).astype(np.float32)
    w2c = torch.inverse(torch.from_numpy(c2w)).numpy()
    for eval_rec in (True, False):
        ref, _ = cull_ref.visibility(special, torch.from_numpy(c2w), torch.from_numpy(g["depths"][:2]), *cam[2:], tr, eval_rec)
        got = helpers.cull_host_frames(special, w2c, g["depths"][:2], cam, tr, eval_rec, 16, 256)
        assert np.array_equal(got.astype(bool), ref), (eval_rec, got, ref)
        assert not got[6:9].any() and not got[0]            # NaN / inf vertices and the camera centre are never seen
    # on the optical axis one metre ahead: inside the frustum, seen without the occlusion test
    assert helpers.cull_host_frames(special[3:4], w2c, g["depths"][:2], cam, tr, False, 16, 256)[0] == 1
    # empty vertex list / no frames / no planes / no faces
    assert helpers.cull_host_frames(np.zeros((0, 3), np.float32), w2c, g["depths"][:2], cam, tr, True, 16, 256).shape == (0,)
    assert not helpers.cull_host_frames(special, w2c[:0], g["depths"][:0], cam, tr, True, 16, 256).any()
    assert helpers.cull_host_hull(special[:6], np.zeros((0, 4), np.float32), 256).all()          # no plane: everything is inside
    nanv = helpers.cull_host_hull(special[6:9], np.array([[1, 0, 0, -1.0]], np.float32), 256)
    assert not nanv[0]                                                                          # NaN is outside (side <= 0 is false)


# ---- the C-ABI from a plain C++ program (no Python, no torch) ------------------------------------------------------------------
def test_abi_consumer_builds_and_its_problem_is_not_trivial():
    """tests/host_harness/cull_gpu_check.cu links against the library through include/unislam_b200.h alone; its synthetic
    problem (host-only mode: the thread functions, no GPU) sees / culls a real fraction of the vertices."""
    import subprocess
    from host_harness import loader
    exe = loader.gpu_check_binary()
    out = subprocess.run([exe, "20000", "8", "host"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    counts = [int(ln.split()[-3]) for ln in out.stdout.strip().splitlines()]
    assert len(counts) == 3 and all(0.05 * 20000 < c < 0.95 * 20000 for c in counts), out.stdout


@pytest.mark.gpu
def test_gpu_abi_consumer_matches_the_host_thread_functions():
    """The same program on the GPU: every culling entry point equals the host build of its thread functions bit for bit, the
    metrics sums agree to 1e-12 (exit code 0), called with cudaMalloc'd pointers from plain C++."""
    import json
    import subprocess
    from host_harness import loader
    out = subprocess.run([loader.gpu_check_binary(), "200000", "24"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["all_equal_host"] is True and res["cull_frames_occlusion"]["mismatch_vs_host"] == 0 and res["render_metrics"]["ok"] is True
