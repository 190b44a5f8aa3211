"""BASELINE-size parity: the CUDA path (MappingStep / TrackingStep through the C-ABI) against the oracle port of the
reference path on the SAME workload and the SAME RNG draws.  TEST INFRASTRUCTURE (imports oracle/): used by
tests/test_gpu_fullsize.py (-m gpu) and by bench.py's cpu_baseline leg, which prints the result as "parity_full_size".

Sizes: BASELINE.json configs[1] (Replica room0: K = 22 frames of P = 81 600 stored pixels, 5982 rays x 40 samples,
decoder variant B, joint_opt) and configs[2] (ScanNet scene0000: 620x460, 56 samples, variant A); the tracker's
2000-ray iteration over the image window that ignore_edge leaves (src/Tracker.py:171-174).

Bars (BASELINE.json north_star): rays / valid mask / sample positions and sample_pdf's searchsorted indices bit-exact;
rendered depth / colour / loss <= 1e-4 relative; decoder, beta, pose and (per level, norm-wise) table gradients
<= 1e-3 relative.  Table gradients are additionally compared with an fp64 run of the oracle, which separates the
summation-order noise of fp32 atomics from real errors.
"""
import importlib
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from oracle import grid_ref, path_ref  # noqa: E402

DEC_NAMES = {
    "A": ["linears.0.weight", "linears.0.bias", "linears.1.weight", "linears.1.bias", "output_linear.weight", "output_linear.bias",
          "c_linears.0.weight", "c_linears.0.bias", "c_linears.1.weight", "c_linears.1.bias", "c_output_linear.weight", "c_output_linear.bias"],
    "B": ["sdf_decoder.params", "color_decoder.params"],
}


def pkg():
    return importlib.import_module("uni-slam_b200")


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1); b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b, floor):
    a = torch.as_tensor(a, dtype=torch.float64); b = torch.as_tensor(b, dtype=torch.float64)
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max()) if a.numel() else 0.0


def trained_like_tables(specs_n, device):
    """Platform-independent 'trained-like' tables (|v| <= 0.05) so that the field, the masks and the gradients are in a
    realistic regime (tcnn's own U(-1e-4,1e-4) init gives a constant field and vanishing table gradients)."""
    return [torch.from_numpy(grid_ref.lcg_params(n, 0.05, 11 + i)).to(device) for i, n in enumerate(specs_n)]


def oracle_field(wl_cpu, tabs, dec, beta, dtype=torch.float32):
    """path_ref.Field over CPU copies of the GPU arm's tensors."""
    cfg = wl_cpu.cfg
    specs = [grid_ref.make_grid_spec(cfg.log2_hash_sdf, wl_cpu.per_level_scale), grid_ref.make_grid_spec(cfg.log2_hash_color, wl_cpu.per_level_scale)]
    w = dict(zip(DEC_NAMES[cfg.decoder_variant], [d.detach().cpu().to(dtype).clone() for d in dec]))
    f = path_ref.Field(specs[0], specs[1], tabs[0].detach().cpu().to(dtype).clone(), tabs[1].detach().cpu().to(dtype).clone(),
                       cfg.decoder_variant, w, beta.detach().cpu().to(dtype).clone(), wl_cpu.bound)
    for t in f.parameters():
        t.requires_grad_(True)
    return f


def to_cpu(wl):
    import copy
    w = copy.copy(wl)
    for k in ("bound", "c2ws", "cam_poses", "depths", "colors", "dirs_cam"):
        setattr(w, k, getattr(wl, k).detach().cpu())
    w.cur_frame = tuple(t.cpu() for t in wl.cur_frame)
    return w


def cpu_draws(wl, gen):
    """The iteration's torch.randint / torch.rand draws, indexed by ray slot (the fused driver's RNG contract)."""
    R, S = wl.n_rays, wl.S
    idx_main = torch.randint(wl.P, (wl.K * wl.n_main,), generator=gen)
    idx_recent = torch.randint(wl.P, (10 * wl.n_recent,), generator=gen) if wl.n_recent else None
    return (idx_main, idx_recent, torch.rand((R, S), generator=gen), torch.rand((R, wl.cfg.n_stratified), generator=gen),
            torch.rand((R, wl.cfg.n_importance), generator=gen))


def oracle_mapping_iteration(wl_cpu, field, draws, joint=True, nodepth_field=None, z_slots_override=None):
    """One reference-path mapping iteration (src/Mapper.py:366-445 up to loss.backward()) through oracle/path_ref on the
    host CPU; draws are slot-indexed and compacted here the way the reference would have drawn them.  Returns a dict
    with the loss, the sampling / rendering intermediates and the gradients.  z_slots_override (R,S): slot-indexed sample
    positions whose depth-less rows replace the oracle's own (path_ref.render_batch_ray, z_nodepth_override)."""
    idx_main, idx_recent, t_rand, t_uni, u_pdf = draws
    cam_poses = wl_cpu.cam_poses.clone().requires_grad_(joint)
    c2ws = torch.cat([wl_cpu.c2ws[0:1], path_ref.cam_pose_to_matrix(cam_poses)], dim=0) if joint else wl_cpu.c2ws
    batches = [(c2ws, wl_cpu.depths, wl_cpu.colors, wl_cpu.dirs_cam, idx_main)]
    if wl_cpu.n_recent:
        K = wl_cpu.K
        batches.append((c2ws[K - 10:], wl_cpu.depths[K - 10:], wl_cpu.colors[K - 10:], wl_cpu.dirs_cam[K - 10:], idx_recent))
    outs = [path_ref.sample_mapping_rays(*b) for b in batches]
    ro = torch.cat([o[0] for o in outs]).detach(); rd = torch.cat([o[1] for o in outs]).detach(); gd = torch.cat([o[2] for o in outs])
    inside = path_ref.bbox_exit(ro, rd, field.bound) >= gd
    has = inside & (gd > 0); holes = inside & ~(gd > 0)
    queue = [t_rand[has]] + ([t_uni[holes], u_pdf[holes]] if holes.any() else [])
    for p in field.parameters():
        p.grad = None
    parts = {}
    cfg = wl_cpu.cfg
    loss = path_ref.mapping_iteration(field, batches, cfg.truncation, cfg.n_stratified, cfg.n_importance,
                                      lambda shape: queue.pop(0), parts=parts, nodepth_field=nodepth_field,
                                      z_nodepth_override=z_slots_override[holes] if (z_slots_override is not None and holes.any()) else None)
    loss.backward()
    return dict(loss=float(loss.detach()), rays_o=ro, rays_d=rd, gt_depth=gd, inside=inside, has=has, holes=holes, parts=parts,
                pose_grad=cam_poses.grad if joint else None, n_inside=int(inside.sum()))


def make_mapping(config_name, device, scale_hw=1.0, seed=1):
    """Workload + field tensors + MappingStep at BASELINE size (scale_hw < 1 only for CPU-sized smoke runs)."""
    P = pkg()
    wlmod = importlib.import_module("uni-slam_b200.workload")
    cfg = P.synthetic.CONFIGS[config_name]
    wl = wlmod.build_mapping_workload(cfg, device, seed=seed, scale_hw=scale_hw)
    meta, tabs, dec, beta = wlmod.init_field_tensors(cfg, wl.bound, wl.per_level_scale, device, seed=0)
    tabs = trained_like_tables([t.numel() for t in tabs], device)
    return wl, meta, tabs, dec, beta


def compare_mapping(step, wl, tabs, dec, beta, draws_cpu, cam_poses, device, fp64_tables=True):
    """Run ONE MappingStep iteration on `draws_cpu` (uploaded) and the oracle on the same draws; return error metrics."""
    wl_cpu = to_cpu(wl)
    wl_cpu.cam_poses = cam_poses.detach().cpu().clone()          # the poses THIS iteration runs on (bench: after its pre-fit)
    R, S = wl.n_rays, wl.S
    dd = [d.to(device) if d is not None else None for d in draws_cpu]
    step.record_pdf_inds(True)
    loss = step.run(wl.batches(dd[0], dd[1]), dd[2], dd[3], dd[4], cam_poses=cam_poses.detach(), c2w_fixed=wl.c2ws[0])
    torch.cuda.synchronize()
    field = oracle_field(wl_cpu, tabs, dec, beta)
    field.tape = [] if fp64_tables else None
    # stage-wise: the depth-less rays' z is compared with the oracle's own resampling (bars below); everything downstream
    # of it (outputs, loss, gradients) is compared along the SAME sample positions -- the inverse-cdf step amplifies ulp-level
    # SDF differences, and an offset of 1e-5 of the range in z would otherwise be charged to the renderer as well
    z_gpu = step.z[:R].cpu()
    o = oracle_mapping_iteration(wl_cpu, field, draws_cpu, z_slots_override=z_gpu)
    ins, has, holes = o["inside"], o["has"], o["holes"]
    ret = o["parts"]["ret"]
    res = {"rays": R, "samples_per_ray": S, "rays_inside": o["n_inside"], "rays_without_depth": int(holes.sum())}
    res["rays_o_mismatch"] = int((step.rays_o[:R].cpu() != o["rays_o"]).sum())
    res["rays_d_mismatch"] = int((step.rays_d[:R].cpu() != o["rays_d"]).sum())
    res["valid_mismatch"] = int((step.valid[:R].cpu().bool() != ins).sum())
    z_ref = torch.zeros_like(z_gpu); z_ref[ins] = ret[5]
    if holes.any():
        z_ref[holes] = o["parts"]["z_nodepth_own"]
    res["z_depth_mismatch"] = int((z_gpu[has] != z_ref[has]).sum())
    res["z_hole_mismatch"] = int((z_gpu[holes] != z_ref[holes]).sum())
    res["z_hole_maxabs"] = float((z_gpu[holes] - z_ref[holes]).abs().max()) if holes.any() else 0.0
    # depth-less rays resample from a cdf of fp32 SDF values (a few ulp apart between the two arithmetics): their sample
    # positions are held to 1e-4 of the ray's sampled range, the searchsorted indices (pdf_inds) to bit-exactness
    res["z_hole_max_rel"] = float(((z_gpu[holes] - z_ref[holes]).abs() / z_ref[holes].amax(dim=1, keepdim=True).clamp_min(1e-3)).max()) if holes.any() else 0.0
    if holes.any():
        res["pdf_inds_mismatch"] = int((step.pdf_inds[:R].cpu()[holes] != o["parts"]["pdf_inds"]).sum())
        res["pdf_inds_checked"] = int(o["parts"]["pdf_inds"].numel())
    else:
        res["pdf_inds_mismatch"], res["pdf_inds_checked"] = 0, 0
    insd = ins.to(device)
    for nm, t, k in (("term", step.term, 0), ("pixel_unc", step.punc, 1), ("depth", step.depth, 2), ("rgb", step.rgb, 3)):
        res[nm + "_rel"] = max_rel(t[:R][insd].cpu(), ret[k].detach(), 1e-3)
    res["sdf_maxabs"] = float((step.raw[:R][insd][..., 3].cpu() - ret[4].detach()).abs().max())
    res["loss_gpu"], res["loss_ref"] = float(loss), o["loss"]
    res["loss_rel"] = abs(float(loss) - o["loss"]) / abs(o["loss"])
    worst = 0.0
    for nm, gt in zip(DEC_NAMES[wl.cfg.decoder_variant], step.fs.g_dec):
        worst = max(worst, rel_err(gt.cpu(), field.w[nm].grad))
    res["dec_grad_rel"] = worst
    res["beta_grad_rel"] = rel_err(step.fs.g_beta.cpu(), field.beta.grad)
    res["pose_grad_rel"] = rel_err(step.d_pose[:wl.K - 1].cpu(), o["pose_grad"])

    def table_errs(ref_tabs, tag):
        for pre, gt, ref, spec in (("sdf", step.fs.g_sdf_table, ref_tabs[0], field.sdf_spec), ("rgb", step.fs.g_rgb_table, ref_tabs[1], field.rgb_spec)):
            g = gt.cpu().double().reshape(-1, 2); r = ref.double().reshape(-1, 2)
            lv_err = [float((g[lv.offset:lv.offset + lv.size] - r[lv.offset:lv.offset + lv.size]).norm()
                            / r[lv.offset:lv.offset + lv.size].norm().clamp_min(1e-30)) for lv in spec.levels]
            res[f"{pre}_table_grad_level_rel_max{tag}"] = max(lv_err)
            res[f"{pre}_table_grad_rel{tag}"] = float((g - r).norm() / r.norm())
            # support: every entry with a non-negligible reference gradient must have been touched.  "Negligible" = below
            # 1e-9 of the table's largest entry: global fp32 atomics flush subnormal contributions (PTX red.add.f32.ftz), and a
            # saturated sigmoid (1 - rgb == 0 in one implementation, 6e-8 in the other) can zero a whole point's tiny share.
            big = r.abs() > 1e-9 * r.abs().max()
            res[f"{pre}_table_support_miss{tag}"] = int(((g == 0) & big).sum())
            res[f"{pre}_table_nnz_excess{tag}"] = int(((g != 0) & (r == 0)).sum())
    table_errs([field.sdf_table.grad, field.rgb_table.grad], "")
    if fp64_tables:
        # the oracle's own d loss / d features, scattered with fp32 weights but accumulated in fp64 (oracle/grid_oracle.c):
        # separates the summation-order noise of fp32 accumulation (GPU atomics and torch's index_add alike) from real errors
        acc = {"sdf": 0, "rgb": 0}
        for gname, pts, h in field.tape:
            spec = field.sdf_spec if gname == "sdf" else field.rgb_spec
            acc[gname] = acc[gname] + grid_ref.c_encode_bwd_params(spec, pts.numpy(), h.grad.numpy())
        ref64 = [torch.from_numpy(np.asarray(acc["sdf"])), torch.from_numpy(np.asarray(acc["rgb"]))]
        table_errs(ref64, "_vs_fp64")
        res["oracle_fp32_vs_fp64_table_grad_rel"] = max(rel_err(field.sdf_table.grad, ref64[0]), rel_err(field.rgb_table.grad, ref64[1]))
    res["table_grad_rel"] = max(res["sdf_table_grad_level_rel_max"], res["rgb_table_grad_level_rel_max"])
    return res


def mapping_ok(r):
    """North-star bars on a compare_mapping result -> list of violated keys (empty = green)."""
    bad = [k for k in ("rays_o_mismatch", "rays_d_mismatch", "valid_mismatch", "z_depth_mismatch", "pdf_inds_mismatch") if r[k] != 0]
    if r["z_hole_max_rel"] >= 1e-4:
        bad.append("z_hole_max_rel")
    bad += [k for k in ("term_rel", "pixel_unc_rel", "depth_rel", "rgb_rel", "loss_rel") if not r[k] < 1e-4]
    bad += [k for k in ("dec_grad_rel", "beta_grad_rel", "pose_grad_rel", "table_grad_rel") if not r[k] < 1e-3]
    for k in r:
        if k.endswith("_vs_fp64") and "rel" in k and not r[k] < 1e-3:
            bad.append(k)
        if "support_miss" in k and r[k] != 0:
            bad.append(k)
    return bad


def run_mapping_fullsize(config_name, device="cuda:0", scale_hw=1.0, fp64_tables=True):
    P = pkg()
    wl, meta, tabs, dec, beta = make_mapping(config_name, device, scale_hw)
    cfg = wl.cfg
    step = P.MappingStep(meta, tabs[0], tabs[1], dec, beta, n_stratified=cfg.n_stratified, n_importance=cfg.n_importance,
                         truncation=cfg.truncation, max_rays=wl.n_rays, max_frames=wl.K)
    draws = cpu_draws(wl, torch.Generator().manual_seed(1234))
    return compare_mapping(step, wl, tabs, dec, beta, draws, wl.cam_poses.clone(), device, fp64_tables)


def run_tracking_fullsize(config_name, device="cuda:0", scale_hw=1.0):
    """TrackingStep at the reference's size (2000 rays over the window ignore_edge leaves, src/Tracker.py:171-174)
    against path_ref.tracking_iteration on the same index / perturbation draws."""
    P = pkg()
    wl, meta, tabs, dec, beta = make_mapping(config_name, device, scale_hw)
    wlmod = importlib.import_module("uni-slam_b200.workload")
    cfg = wl.cfg
    col, dep, c2w = wl.cur_frame
    H, W = dep.shape
    sc = W / cfg.cam.W
    fx, fy, cx, cy = cfg.cam.fx * sc, cfg.cam.fy * sc, cfg.cam.cx * sc, cfg.cam.cy * sc
    e = int(cfg.ignore_edge * sc)
    n = cfg.track_pixels
    trk = P.TrackingStep(meta, tabs[0], tabs[1], dec, beta, n_stratified=cfg.n_stratified, n_importance=cfg.n_importance,
                         truncation=cfg.truncation, H=H, W=W, fx=fx, fy=fy, cx=cx, cy=cy, ignore_edge_h=e, ignore_edge_w=e, n_rays=n)
    pose = wlmod._matrix_to_cam_pose(c2w[None]).contiguous()
    pose[:, 4:] += torch.tensor([0.012, -0.008, 0.005], device=device)
    pose[:, :4] += torch.tensor([0.002, -0.003, 0.001, 0.002], device=device)       # un-normalised quaternion on purpose
    gen = torch.Generator().manual_seed(4321)
    idx = torch.randint((H - 2 * e) * (W - 2 * e), (n,), generator=gen)
    t_rand_slot = torch.rand((n, trk.S), generator=gen)
    loss = trk.run(pose, dep, col, idx.to(device), t_rand_slot.to(device))
    torch.cuda.synchronize()
    wl_cpu = to_cpu(wl)
    field = oracle_field(wl_cpu, tabs, dec, beta)
    for t in field.parameters():
        t.requires_grad_(False)
    cam_pose = pose.detach().cpu().clone().requires_grad_(True)
    # the reference draws (n_valid, S): compact the slot-indexed draw with the oracle's own validity mask
    with torch.no_grad():
        c2w_ref = path_ref.cam_pose_to_matrix(cam_pose)
        ro, rd, gd, _ = path_ref.sample_tracking_rays(e, H - e, e, W - e, n, fx, fy, cx, cy, c2w_ref, dep.cpu()[None], col.cpu()[None], idx)
        inside = (path_ref.bbox_exit(ro, rd, field.bound) >= gd) & (gd > 0)
    parts = {}
    queue = [t_rand_slot[inside]]
    loss_ref, _ = path_ref.tracking_iteration(field, cam_pose, dep.cpu()[None], col.cpu()[None], H, W, fx, fy, cx, cy, e, e, idx,
                                              cfg.truncation, cfg.n_stratified, cfg.n_importance, lambda shape: queue.pop(0), parts=parts)
    loss_ref.backward()
    ret = parts["ret"]
    res = {"rays": n, "rays_valid": int(inside.sum()), "window": [H - 2 * e, W - 2 * e]}
    res["rays_o_mismatch"] = int((trk.rays_o.cpu() != ro).sum())
    res["rays_d_mismatch"] = int((trk.rays_d.cpu() != rd).sum())
    res["valid_mismatch"] = int((trk.valid.cpu().bool() != inside).sum())
    res["z_mismatch"] = int((trk.z.cpu()[inside] != ret[5]).sum())
    insd = inside.to(device)
    for nm, t, k in (("term", trk.term, 0), ("pixel_unc", trk.punc, 1), ("depth", trk.depth, 2), ("rgb", trk.rgb, 3)):
        res[nm + "_rel"] = max_rel(t[insd].cpu(), ret[k].detach(), 1e-3)
    res["median_mismatch"] = int(float(trk.median) != float(parts["median"]))
    res["loss_gpu"], res["loss_ref"] = float(loss), float(loss_ref.detach())
    res["loss_rel"] = abs(float(loss) - float(loss_ref.detach())) / abs(float(loss_ref.detach()))
    res["grad_T_rel"] = rel_err(trk.d_pose[:, 4:].cpu(), cam_pose.grad[:, 4:])
    res["grad_R_rel"] = rel_err(trk.d_pose[:, :4].cpu(), cam_pose.grad[:, :4])
    return res


def oracle_tracking_setup(wl, tabs, dec, beta, device_pose_offset=True):
    """CPU oracle field + one tracking problem (pose, window, draws) at the reference's size, for bench.py's CPU baseline."""
    P = pkg()
    wlmod = importlib.import_module("uni-slam_b200.workload")
    wl_cpu = to_cpu(wl)
    cfg = wl.cfg
    col, dep, c2w = wl_cpu.cur_frame
    H, W = dep.shape
    sc = W / cfg.cam.W
    cam = (H, W, cfg.cam.fx * sc, cfg.cam.fy * sc, cfg.cam.cx * sc, cfg.cam.cy * sc)
    e = int(cfg.ignore_edge * sc)
    field = oracle_field(wl_cpu, tabs, dec, beta)
    for t in field.parameters():
        t.requires_grad_(False)
    pose = wlmod._matrix_to_cam_pose(c2w[None]).contiguous()
    pose[:, 4:] += torch.tensor([0.012, -0.008, 0.005])
    return wl_cpu, field, pose, cam, e


def oracle_tracking_iteration(wl_cpu, field, pose, cam, e, gen):
    """One Tracker.optimize_tracking iteration (src/Tracker.py:149-244) through oracle/path_ref on the host CPU: draw, forward,
    loss, pose gradient."""
    cfg = wl_cpu.cfg
    H, W, fx, fy, cx, cy = cam
    col, dep, _ = wl_cpu.cur_frame
    n = cfg.track_pixels
    S = cfg.n_stratified + cfg.n_importance
    idx = torch.randint((H - 2 * e) * (W - 2 * e), (n,), generator=gen)
    cam_pose = pose.clone().requires_grad_(True)
    draw = lambda shape: torch.rand(shape, generator=gen)
    loss, _ = path_ref.tracking_iteration(field, cam_pose, dep[None], col[None], H, W, fx, fy, cx, cy, e, e, idx, cfg.truncation,
                                          cfg.n_stratified, cfg.n_importance, draw)
    loss.backward()
    return float(loss.detach()), cam_pose.grad


def oracle_dense_query_chunk(wl_cpu, field, n_points=500000):
    """Mesher.eval_points (src/utils/Mesher.py:134-166) on one points_batch_size chunk of the 1 cm query grid, SDF channel,
    through oracle/path_ref on the host CPU.  Returns the number of points evaluated."""
    axes = path_ref.mesh_grid_axes([[lo, hi] for lo, hi in wl_cpu.cfg.bound_yaml], resolution=0.01)
    nx, ny, nz = [a.numel() for a in axes]
    # a chunk of consecutive points in the reference's flattening order (meshgrid 'xy': y slowest, then x, then z)
    m = min(n_points, nx * ny * nz)
    flat = torch.arange(m)
    iz = flat % nz; ix = (flat // nz) % nx; iy = flat // (nz * nx)
    pts = torch.stack([axes[0][ix], axes[1][iy + ny // 2], axes[2][iz]], dim=1)
    with torch.no_grad():
        path_ref.eval_points_sdf(field, pts)
    return m


def tracking_ok(r):
    bad = [k for k in ("rays_o_mismatch", "rays_d_mismatch", "valid_mismatch", "z_mismatch") if r[k] != 0]
    bad += [k for k in ("term_rel", "pixel_unc_rel", "depth_rel", "rgb_rel", "loss_rel") if not r[k] < 1e-4]
    bad += [k for k in ("grad_T_rel", "grad_R_rel") if not r[k] < 1e-3]
    return bad


if __name__ == "__main__":
    import json
    name = sys.argv[1] if len(sys.argv) > 1 else "replica_room0"
    r = run_mapping_fullsize(name)
    print(json.dumps(r, indent=1)); print("violations:", mapping_ok(r))
    r = run_tracking_fullsize(name)
    print(json.dumps(r, indent=1)); print("violations:", tracking_ok(r))
