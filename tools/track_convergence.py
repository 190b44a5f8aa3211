"""End-to-end sanity of the pose gradient: fit the field on a keyframe window, then track a HELD-OUT frame from a
perturbed pose and report the translation / rotation error per iteration (should shrink towards the analytic GT)."""
import importlib, json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
P = importlib.import_module("uni-slam_b200")
wlmod = importlib.import_module("uni-slam_b200.workload")
syn = P.synthetic
dev = "cuda:0"
torch.manual_seed(0)
cfg = syn.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "replica_room0"]
scale = 0.5
wl = wlmod.build_mapping_workload(cfg, dev, seed=1, scale_hw=scale)
wl.c2ws = wl.c2ws.clone()
meta, tabs, dec, beta = wlmod.init_field_tensors(cfg, wl.bound, wl.per_level_scale, dev)
step = P.MappingStep(meta, tabs[0], tabs[1], dec, beta, n_stratified=cfg.n_stratified, n_importance=cfg.n_importance,
                     truncation=cfg.truncation, max_rays=wl.n_rays, max_frames=wl.K)
params = [tabs[0], tabs[1], beta] + dec
for p_, g_ in zip(params, [step.fs.g_sdf_table, step.fs.g_rgb_table, step.fs.g_beta] + step.fs.g_dec):
    p_.requires_grad_(True); p_.grad = g_
opt = P.FusedAdam([{"params": dec + [beta], "lr": 1e-3}, {"params": [tabs[0]], "lr": cfg.hash_lr}, {"params": [tabs[1]], "lr": cfg.hash_lr}])
losses = []
for it in range(400):
    idx_main, idx_recent, t_rand, t_uni, u_pdf = wl.draw()
    step.run(wl.batches(idx_main, idx_recent), t_rand, t_uni, u_pdf)
    opt.step()
    if it % 100 == 0 or it == 399: losses.append(round(float(step.loss), 4))
# held-out frame between two keyframes
seq = syn.SyntheticSequence(cfg, n_frames=200, device=dev, seed=1, scale_hw=scale)
k = (wl.K - 2) * cfg.map_every + cfg.map_every // 2
col, dep, c2w_gt = seq.frame(k)
cam = seq.cam
e = max(int(cfg.ignore_edge * scale), 2)
trk = P.TrackingStep(meta, tabs[0].detach(), tabs[1].detach(), [d.detach() for d in dec], beta.detach(), n_stratified=cfg.n_stratified,
                     n_importance=cfg.n_importance, truncation=cfg.truncation, H=cam.H, W=cam.W, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy,
                     ignore_edge_h=e, ignore_edge_w=e, n_rays=cfg.track_pixels)
pose_gt = wlmod._matrix_to_cam_pose(c2w_gt[None])
pose = pose_gt.clone()
pose[:, 4:] += torch.tensor([0.02, -0.015, 0.01], device=dev)
ang = math.radians(1.0)
dq = torch.tensor([math.cos(ang / 2), math.sin(ang / 2), 0, 0], device=dev)
q = pose[0, :4]
pose[0, :4] = torch.stack([dq[0]*q[0]-dq[1]*q[1], dq[0]*q[1]+dq[1]*q[0], dq[0]*q[2]-dq[1]*q[3], dq[0]*q[3]+dq[1]*q[2]])
T_ = pose[:, 4:].clone().contiguous().requires_grad_(True); R_ = pose[:, :4].clone().contiguous().requires_grad_(True)
T_.grad = trk.d_pose[:, 4:]; R_.grad = trk.d_pose[:, :4]
topt = P.FusedAdam([{"params": [T_], "lr": cfg.lr_T, "betas": (0.5, 0.999)}, {"params": [R_], "lr": cfg.lr_R, "betas": (0.5, 0.999)}])
npx = (cam.H - 2 * e) * (cam.W - 2 * e)
def errs():
    t_err = float((T_.detach() - pose_gt[:, 4:]).norm())
    qa = R_.detach()[0] / R_.detach()[0].norm(); qb = pose_gt[0, :4] / pose_gt[0, :4].norm()
    r_err = math.degrees(2 * math.acos(min(1.0, abs(float((qa * qb).sum())))))
    return round(t_err * 100, 3), round(r_err, 3)
hist = [errs()]
for it in range(60):
    cam_pose = torch.cat([R_.detach(), T_.detach()], -1).contiguous()
    trk.run(cam_pose, dep, col, torch.randint(npx, (cfg.track_pixels,), device=dev), torch.rand((cfg.track_pixels, trk.S), device=dev))
    topt.step()
    if it % 10 == 9: hist.append(errs())
print(json.dumps({"config": cfg.name, "map_losses": losses, "track_err_cm_deg_every10": hist, "track_loss_end": float(trk.loss)}))
