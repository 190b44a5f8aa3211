"""f4, the culling after marching cubes (src/tools/cull_mesh.py): csrc/cull.cu + mesh.MeshCuller.

Chain of evidence:
  unmodified cull_mesh / cull_out_bound_mesh (run by oracle/gen_golden.py) -> tests/golden/cull_replica.npz
  oracle/cull_ref.py == golden                                                    (CPU)
  element functions of the kernels (csrc/usl_cull.cuh, compiled for the host by tests/host_harness) == golden, and == oracle on
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
    for fpc in (16, 3, 1):
        seen = helpers.cull_host_frames(g["verts"], w2c, g["depths"], cam, tr, eval_rec, fpc)
        assert np.array_equal(seen.astype(bool), g["seen_" + tag]), fpc
    col = (np.arange(len(g["verts"]) * 3) % 251).astype(np.uint8).reshape(-1, 3)
    v, f, c, keep = helpers.cull_host_compact(g["verts"], col, g["faces"], seen, require_all=0)
    assert np.array_equal(keep.astype(bool), g["face_keep_" + tag])
    assert np.array_equal(v, g["culled_verts_" + tag]) and np.array_equal(f, g["culled_faces_" + tag])
    _, _, c_ref, _ = cull_ref.face_filter(g["verts"], g["faces"], seen.astype(bool), False, colors=col)
    assert np.array_equal(c, c_ref)
    planes = pkg().mesh.hull_planes(g["hull_verts"], g["hull_faces"])
    inside = helpers.cull_host_hull(g["verts"], planes)
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
    # accumulation over ranges of frames == one call
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    acc = culler.seen_by_frames(t(pts), t(c2ws[:20]), t(depths[:20]), eval_rec)
    acc = culler.seen_by_frames(t(pts), t(c2ws[20:]), t(depths[20:]), eval_rec, seen=acc)
    assert np.array_equal(acc.cpu().numpy(), ref)
    seen_ref, margin = cull_ref.visibility(pts, torch.from_numpy(c2ws), torch.from_numpy(depths), *cam[2:], tr, eval_rec)
    diff = seen.astype(bool) != seen_ref
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
