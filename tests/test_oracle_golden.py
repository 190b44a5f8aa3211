"""CPU: oracle/path_ref.py (the restatement) against fixtures produced by the UNMODIFIED reference
(oracle/gen_golden.py). This is what pins the oracle for everything but the tcnn arithmetic."""
import numpy as np
import pytest
import torch

from oracle import path_ref
from helpers import DrawQueue, golden_field, load_golden, max_rel, rel_err

T = torch.from_numpy


def _draws(g):
    d = [T(g["t_rand"])]
    if "t_rand_uni" in g:
        d += [T(g["t_rand_uni"]), T(g["u_pdf"])]
    return DrawQueue(d)


@pytest.mark.parametrize("name", ["map_replica_k1", "map_replica_k7", "map_scannet_k23", "map_replica_nomask", "map_replica_kfstore"])
def test_mapping_iteration_matches_reference(name):
    g = load_golden(name)
    field = golden_field(g, 0)
    joint = int(g["joint_opt"]) == 1
    c2ws = T(g["call0_c2ws"])
    if joint:
        cam_poses = T(g["cam_poses"]).clone().requires_grad_(True)
        c2ws = torch.cat([c2ws[0:1], path_ref.cam_pose_to_matrix(cam_poses)], dim=0)
        assert torch.equal(c2ws.detach(), T(g["call0_c2ws"]))
    batches = []
    for ci in range(int(g["n_sample_calls"])):
        cw = c2ws if ci == 0 else c2ws[-10:]
        batches.append((cw, T(g[f"call{ci}_depths"]), T(g[f"call{ci}_colors"]), T(g[f"call{ci}_rays_d_cam"]),
                        T(g[f"call{ci}_indices"])))
        o = path_ref.sample_mapping_rays(*batches[-1])
        for nm, t in zip(("rays_o", "rays_d", "depth", "color"), o):
            assert torch.equal(t.detach(), T(g[f"call{ci}_out_{nm}"])), nm       # bit-exact sampling
    parts = {}
    loss = path_ref.mapping_iteration(field, batches, float(g["truncation"]), int(g["n_stratified"]),
                                      int(g["n_importance"]), _draws(g), parts=parts,
                                      mask_mode=str(g["mask_mode"]) if "mask_mode" in g else "original")
    assert torch.equal(parts["rays_o"].detach(), T(g["render_rays_o"]))
    assert torch.equal(parts["gt_depth"], T(g["render_gt_depth"]))
    ret = parts["ret"]
    assert torch.equal(ret[5], T(g["ret_z_vals"]))                               # z_vals bit-exact (incl. sample_pdf path)
    if "pdf_inds" in g:                                                          # searchsorted indices of sample_pdf (common.py:70)
        assert torch.equal(parts["pdf_inds"], T(g["pdf_inds"]))
    for nm, t in zip(("term", "pixel_unc", "depth", "rgb", "sdf"), ret[:5]):
        assert max_rel(t.detach(), g["ret_" + nm], 1e-4) < 1e-5, nm
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-6
    loss.backward()
    for k in g:
        if k.startswith("grad_dec."):
            nm = k[len("grad_dec."):]
            got = field.beta.grad if nm == "beta" else field.w[nm].grad
            assert rel_err(got, g[k]) < 1e-5, nm
    for pre, tab, spec in (("grad_sdf_table", field.sdf_table, field.sdf_spec), ("grad_rgb_table", field.rgb_table, field.rgb_spec)):
        gr = tab.grad.reshape(-1)
        assert int((gr != 0).sum()) == int(g[pre + "_nnz"])
        assert rel_err(gr[T(g[pre + "_idx"])], g[pre + "_val"]) < 1e-5
        norms = [gr.double().reshape(-1, 2)[lv.offset:lv.offset + lv.size].norm().item() for lv in spec.levels]
        np.testing.assert_allclose(norms, g[pre + "_level_norm"], rtol=1e-5)
    if joint:
        assert rel_err(cam_poses.grad, g["grad_cam_poses"]) < 1e-5


def test_nodepth_z_override_is_what_the_renderer_integrates_along():
    """The stage-wise full-size gate (tests/fullsize_cases.py) feeds the checked implementation's depth-less sample positions
    to the oracle: the override must replace exactly those rows, leave the oracle's own resampling on record, and reproduce
    the unmodified result when it equals it."""
    g = load_golden("map_replica_k7")
    assert "u_pdf" in g                                            # the case has depth-less rays
    c2ws = T(g["call0_c2ws"])
    batches = [(c2ws, T(g["call0_depths"]), T(g["call0_colors"]), T(g["call0_rays_d_cam"]), T(g["call0_indices"]))]
    args = (float(g["truncation"]), int(g["n_stratified"]), int(g["n_importance"]))

    def run(override):
        parts = {}
        loss = path_ref.mapping_iteration(golden_field(g, 0), batches, *args, _draws(g), parts=parts, z_nodepth_override=override)
        return float(loss), parts
    loss0, p0 = run(None)
    own = p0["z_nodepth_own"]
    holes = ~(p0["gt_depth"] > 0)
    assert torch.equal(p0["ret"][5][holes], own) and int(holes.sum()) == own.shape[0] > 0
    loss1, p1 = run(own.clone())                                   # override == own z: nothing changes
    assert loss1 == loss0 and torch.equal(p1["ret"][5], p0["ret"][5])
    shifted = own + 1e-3
    loss2, p2 = run(shifted)
    assert torch.equal(p2["ret"][5][holes], shifted) and torch.equal(p2["ret"][5][~holes], p0["ret"][5][~holes])
    assert torch.equal(p2["z_nodepth_own"], own) and torch.equal(p2["pdf_inds"], p0["pdf_inds"])
    assert loss2 != loss0


@pytest.mark.parametrize("name", ["track_replica", "track_scannet", "track_scannet_nomask"])
def test_tracking_iteration_matches_reference(name):
    g = load_golden(name)
    field = golden_field(g, 50, requires_grad=False)
    H, W, fx, fy, cx, cy = [float(v) for v in g["meta_H_W_fx_fy_cx_cy"]]
    H, W = int(H), int(W)
    cam_pose = T(g["cam_pose"]).clone().requires_grad_(True)
    e = int(g["edge"])
    parts = {}
    loss, punc = path_ref.tracking_iteration(field, cam_pose, T(g["depth_img"])[None], T(g["color_img"])[None],
                                             H, W, fx, fy, cx, cy, e, e, T(g["indices"]), float(g["truncation"]),
                                             int(g["n_stratified"]), int(g["n_importance"]), DrawQueue([T(g["t_rand"])]),
                                             parts=parts, mask_mode=str(g["mask_mode"]) if "mask_mode" in g else "original")
    assert torch.equal(parts["rays_o"].detach(), T(g["render_rays_o"]))
    assert torch.equal(parts["rays_d"].detach(), T(g["render_rays_d"]))
    assert torch.equal(parts["ret"][5], T(g["ret_z_vals"]))
    for nm, t in zip(("term", "pixel_unc", "depth", "rgb", "sdf"), parts["ret"][:5]):
        assert max_rel(t.detach(), g["ret_" + nm], 1e-4) < 1e-5, nm
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-6
    loss.backward()
    assert rel_err(cam_pose.grad[:, 4:], g["grad_T"]) < 1e-5
    assert rel_err(cam_pose.grad[:, :4], g["grad_R"]) < 1e-5


@pytest.mark.parametrize("name", ["img_replica", "img_scannet"])
def test_render_img_matches_reference(name):
    """Renderer.render_img (Renderer.py:160-223) of the unmodified reference vs oracle/path_ref.render_img."""
    from helpers import golden_img_draws
    g = load_golden(name)
    field = golden_field(g, 80, requires_grad=False)
    H, W, fx, fy, cx, cy = g["meta_H_W_fx_fy_cx_cy"]
    H, W = int(H), int(W)
    ret = path_ref.render_img(field, H, W, fx, fy, cx, cy, T(g["c2w"]), T(g["depth_img"]), int(g["n_stratified"]),
                              int(g["n_importance"]), float(g["truncation"]), int(g["ray_batch"]), DrawQueue(golden_img_draws(g)))
    for nm, t in zip(("depth", "color", "term", "pixel_unc", "depth_unc"), ret):
        assert str(t.dtype) == str(g["dtype_" + nm]), nm                          # float64 except colour (Renderer.py:205-209)
        assert tuple(t.shape) == g["ret_" + nm].shape
        assert max_rel(t, g["ret_" + nm], 1e-4) < 1e-5, nm


@pytest.mark.parametrize("name", ["mesh_replica", "mesh_scannet"])
def test_mesh_query_matches_reference(name):
    """Mesher.get_grid_uniform + eval_points (Mesher.py:134-195) of the unmodified reference vs the oracle's restatement:
    axis sample counts and coordinates, point ordering of meshgrid(indexing='xy'), strict in-bound mask, SDF values."""
    g = load_golden(name)
    field = golden_field(g, 120, requires_grad=False)
    axes = path_ref.mesh_grid_axes(g["mc_bound"], resolution=float(g["resolution"]))
    for a, nm in zip(axes, "xyz"):
        assert torch.equal(a, T(g["axis_" + nm]).float()), nm                    # counts and fp32 coordinates bit-exact
    pts = path_ref.mesh_grid_points(axes)
    assert torch.equal(pts, T(g["points"]))
    with torch.no_grad():
        sdf = path_ref.eval_points_sdf(field, pts)
    ref = T(g["sdf"])
    assert torch.equal(sdf == -1, ref == -1)
    assert (sdf - ref).abs().max() < 1e-6


def test_keyframe_covisibility_matches_reference():
    """Mapper.keyframe_selection_LC's overlap measure (Mapper.py:177-236) of the unmodified reference vs the oracle's restatement."""
    g = load_golden("kf_covis_replica")
    H, W, fx, fy, cx, cy = [float(v) for v in g["meta_H_W_fx_fy_cx_cy"]]
    ro, rd, d, _ = path_ref.sample_tracking_rays(0, int(H), 0, int(W), int(g["num_rays"]), fx, fy, cx, cy, T(g["c2w"])[None],
                                                 T(g["depth_img"])[None], T(g["color_img"])[None], T(g["indices"]))
    assert torch.equal(ro, T(g["sample_out_rays_o"])) and torch.equal(rd, T(g["sample_out_rays_d"])) and torch.equal(d, T(g["sample_out_depth"]))
    pi = path_ref.keyframe_covisibility(ro, rd, d, T(g["keyframe_c2ws"])[:-2], int(H), int(W), fx, fy, cx, cy, int(g["num_samples"]), int(g["edge"]))
    assert torch.equal(pi, T(g["percent_inside"]))
