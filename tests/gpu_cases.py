"""Runs golden cases (tests/golden/*.npz, produced by the UNMODIFIED reference) through the CUDA path
and returns error metrics.  Used by the -m gpu tests and by __graft_entry__.smoke()."""
import numpy as np
import torch

from oracle import grid_ref, path_ref
from helpers import DEC_SHAPES, golden_decoder_weights, load_golden, pkg, rel_err, max_rel

T = torch.from_numpy

DEC_ORDER = {
    "A": ["linears.0.weight", "linears.0.bias", "linears.1.weight", "linears.1.bias", "output_linear.weight", "output_linear.bias",
          "c_linears.0.weight", "c_linears.0.bias", "c_linears.1.weight", "c_linears.1.bias", "c_output_linear.weight", "c_output_linear.bias"],
    "B": ["sdf_decoder.params", "color_decoder.params"],
}


def cuda_field(g, seed_salt, device):
    """FieldMeta + device tensors with the same integer-hash fill as the golden generator."""
    P = pkg()
    variant = str(g["variant"])
    grids = [P._lib.build_grid(16, int(g["log2_hash"][i]), 16, float(g["per_level_scale"][i])) for i in range(2)]
    tabs = [T(grid_ref.lcg_params(grids[i].total_entries * 2, 0.05, i + 1 + seed_salt)).to(device) for i in range(2)]
    w = golden_decoder_weights(variant, seed_salt)
    dec = [w[k].to(device).contiguous() for k in DEC_ORDER[variant]]
    beta = torch.full((1,), 10.0, device=device)
    meta = P.ops.FieldMeta(grids[0], grids[1], variant, P._lib.make_bound(T(g["bound"])))
    return meta, tabs, dec, beta


def _slot_draws(g, inside, gt_all, S, ns, ni, device):
    """Scatter the reference's compacted torch.rand draws into per-ray-slot tensors."""
    R = gt_all.shape[0]
    t_rand = torch.zeros((R, S)); t_uni = torch.zeros((R, ns)); u_pdf = torch.zeros((R, ni))
    has = inside & (gt_all > 0); holes = inside & ~(gt_all > 0)
    t_rand[has] = T(g["t_rand"])
    if holes.any():
        t_uni[holes] = T(g["t_rand_uni"]); u_pdf[holes] = T(g["u_pdf"])
    return t_rand.to(device), t_uni.to(device), u_pdf.to(device), bool(holes.any())


def run_mapping_case(name, device="cuda:0"):
    P = pkg()
    g = load_golden(name)
    meta, tabs, dec, beta = cuda_field(g, 0, device)
    ns, ni = int(g["n_stratified"]), int(g["n_importance"]); S = ns + ni
    tr = float(g["truncation"])
    joint = int(g["joint_opt"]) == 1
    ncall = int(g["n_sample_calls"])
    K = g["call0_c2ws"].shape[0]
    batches, depth_all, ro_all, rd_all = [], [], [], []
    for ci in range(ncall):
        Kb = g[f"call{ci}_depths"].shape[0]
        batches.append((T(g[f"call{ci}_c2ws"]).to(device).contiguous(), T(g[f"call{ci}_depths"]).to(device), T(g[f"call{ci}_colors"]).to(device),
                        T(g[f"call{ci}_rays_d_cam"]).to(device), T(g[f"call{ci}_indices"]).to(device), int(g[f"call{ci}_n"]), K - Kb))
        depth_all.append(T(g[f"call{ci}_out_depth"])); ro_all.append(T(g[f"call{ci}_out_rays_o"])); rd_all.append(T(g[f"call{ci}_out_rays_d"]))
    depth_all, ro_all, rd_all = torch.cat(depth_all), torch.cat(ro_all), torch.cat(rd_all)
    R = depth_all.shape[0]
    inside = path_ref.bbox_exit(ro_all, rd_all, T(g["bound"])) >= depth_all
    t_rand, t_uni, u_pdf, has_holes = _slot_draws(g, inside, depth_all, S, ns, ni, device)
    step = P.MappingStep(meta, tabs[0], tabs[1], dec, beta, n_stratified=ns, n_importance=ni, truncation=tr,
                         max_rays=R, max_frames=K, mask_mode=str(g["mask_mode"]) if "mask_mode" in g else "original")
    cam_poses = T(g["cam_poses"]).to(device).contiguous() if joint else None
    step.record_pdf_inds(True)
    loss = step.run(batches, t_rand, t_uni, u_pdf, cam_poses=cam_poses,
                    c2w_fixed=T(g["call0_c2ws"][0]).to(device) if joint else None, has_holes=has_holes)
    torch.cuda.synchronize()
    res = {}
    ins = inside.to(device)
    # sampling parity (bit-exact)
    res["rays_o_mismatch"] = float((step.rays_o[:R].cpu() != ro_all).sum())
    res["rays_d_mismatch"] = float((step.rays_d[:R].cpu() != rd_all).sum())
    res["valid_mismatch"] = float((step.valid[:R].cpu().bool() != inside).sum())
    zg = T(g["ret_z_vals"])
    zc = step.z[:R][ins].cpu()
    has_depth = T(g["render_gt_depth"]) > 0
    res["z_depth_mismatch"] = float((zc[has_depth] != zg[has_depth]).sum())          # depth-guided rays: bit-exact
    res["z_hole_maxabs"] = float((zc[~has_depth] - zg[~has_depth]).abs().max()) if (~has_depth).any() else 0.0
    # "sample indices bit-exact": torch.searchsorted of sample_pdf (common.py:70) as the unmodified reference produced them
    res["pdf_inds_mismatch"] = float((step.pdf_inds[:R][ins].cpu()[~has_depth] != T(g["pdf_inds"])).sum()) if "pdf_inds" in g else 0.0
    for nm, t in (("term", step.term), ("pixel_unc", step.punc), ("depth", step.depth), ("rgb", step.rgb)):
        res[nm + "_rel"] = max_rel(t[:R][ins].cpu(), g["ret_" + nm], 1e-3)
    res["sdf_rel"] = max_rel(step.raw[:R][ins][..., 3].cpu(), g["ret_sdf"], 1e-2)
    res["loss_rel"] = abs(float(loss) - float(g["loss"])) / abs(float(g["loss"]))
    # gradients
    variant = str(g["variant"])
    worst = 0.0
    for k, gt in zip(DEC_ORDER[variant], step.fs.g_dec):
        worst = max(worst, rel_err(gt.cpu(), g["grad_dec." + k]))
    res["dec_grad_rel"] = worst
    res["beta_grad_rel"] = rel_err(step.fs.g_beta.cpu(), g["grad_dec.beta"])
    worst = 0.0
    for pre, gt in (("grad_sdf_table", step.fs.g_sdf_table), ("grad_rgb_table", step.fs.g_rgb_table)):
        gc = gt.cpu().reshape(-1)
        worst = max(worst, rel_err(gc[T(g[pre + "_idx"])], g[pre + "_val"]))
        spec = grid_ref.make_grid_spec(int(g["log2_hash"][0 if "sdf" in pre else 1]), float(g["per_level_scale"][0 if "sdf" in pre else 1]))
        norms = np.array([gc.double().reshape(-1, 2)[lv.offset:lv.offset + lv.size].norm().item() for lv in spec.levels])
        worst = max(worst, float(np.max(np.abs(norms - g[pre + "_level_norm"]) / np.maximum(g[pre + "_level_norm"], 1e-12))))
        # global fp32 atomics flush subnormal contributions to zero (PTX red.add.f32), torch's CPU index_add keeps
        # them: compare the support above the subnormal range instead of the raw non-zero count
        big = np.abs(g[pre + "_val"]) > 1e-30
        res[pre + "_support_miss"] = float((gc[T(g[pre + "_idx"])][T(big)] == 0).sum())
        res[pre + "_nnz_excess"] = float(max(int((gc != 0).sum()) - int(g[pre + "_nnz"]), 0))
    res["table_grad_rel"] = worst
    res["pose_grad_rel"] = rel_err(step.d_pose[:K - 1].cpu(), g["grad_cam_poses"]) if joint else 0.0
    return res


def run_tracking_case(name, device="cuda:0"):
    P = pkg()
    g = load_golden(name)
    meta, tabs, dec, beta = cuda_field(g, 50, device)
    ns, ni = int(g["n_stratified"]), int(g["n_importance"]); S = ns + ni
    H, W, fx, fy, cx, cy = [float(v) for v in g["meta_H_W_fx_fy_cx_cy"]]
    H, W = int(H), int(W)
    e = int(g["edge"])
    idx = T(g["indices"])
    R = idx.shape[0]
    gt_all = T(g["sample_out_depth"])
    inside = (path_ref.bbox_exit(T(g["sample_out_rays_o"]), T(g["sample_out_rays_d"]), T(g["bound"])) >= gt_all) & (gt_all > 0)
    t_rand = torch.zeros((R, S)); t_rand[inside] = T(g["t_rand"])
    step = P.TrackingStep(meta, tabs[0], tabs[1], dec, beta, n_stratified=ns, n_importance=ni, truncation=float(g["truncation"]),
                          H=H, W=W, fx=fx, fy=fy, cx=cx, cy=cy, ignore_edge_h=e, ignore_edge_w=e, n_rays=R,
                          mask_mode=str(g["mask_mode"]) if "mask_mode" in g else "original")
    cam_pose = T(g["cam_pose"]).to(device).contiguous()
    loss = step.run(cam_pose, T(g["depth_img"]).to(device).contiguous(), T(g["color_img"]).to(device).contiguous(),
                    idx.to(device), t_rand.to(device))
    torch.cuda.synchronize()
    ins = inside.to(device)
    res = {}
    res["rays_o_mismatch"] = float((step.rays_o.cpu() != T(g["sample_out_rays_o"])).sum())
    res["rays_d_mismatch"] = float((step.rays_d.cpu() != T(g["sample_out_rays_d"])).sum())
    res["valid_mismatch"] = float((step.valid.cpu().bool() != inside).sum())
    res["z_mismatch"] = float((step.z[ins].cpu() != T(g["ret_z_vals"])).sum())
    for nm, t in (("term", step.term), ("pixel_unc", step.punc), ("depth", step.depth), ("rgb", step.rgb)):
        res[nm + "_rel"] = max_rel(t[ins].cpu(), g["ret_" + nm], 1e-3)
    res["loss_rel"] = abs(float(loss) - float(g["loss"])) / abs(float(g["loss"]))
    res["grad_T_rel"] = rel_err(step.d_pose[:, 4:].cpu(), g["grad_T"])
    res["grad_R_rel"] = rel_err(step.d_pose[:, :4].cpu(), g["grad_R"])
    # (1-term)^2 with term ~ 1 is ill-conditioned in fp32: compare with an absolute floor of 1 ulp(term)^2-ish
    ref = float(g["pixel_unc"].mean())
    res["mean_punc_err"] = abs(float(step.acc[11] / step.acc[9]) - ref) / (ref + 1e-6)
    return res
