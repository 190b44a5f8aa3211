"""Profiling target for ncu: one pass each of the paths outside the mapping step -- the dense SDF query (config 5), marching
cubes on its volume, the whole-frame renderer (render_img) and one tracking iteration -- on the Replica-shaped field.
  ncu ... python tools/profile_targets.py"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
P = importlib.import_module("uni-slam_b200")
wlmod = importlib.import_module("uni-slam_b200.workload")
meshmod = importlib.import_module("uni-slam_b200.mesh")
dev = "cuda:0"
cfg = P.synthetic.REPLICA_ROOM0
wl = wlmod.build_mapping_workload(cfg, dev, seed=1, n_keyframes=2)
meta, tabs, dec, beta = wlmod.init_field_tensors(cfg, wl.bound, wl.per_level_scale, dev, seed=0)
tabs = [torch.randn_like(t) * 0.05 for t in tabs]
axes = []
for a in range(3):
    lo, hi = cfg.bound_yaml[a]
    axes.append(torch.from_numpy(np.linspace(lo - 0.05, hi + 0.05, int(round((hi - lo + 0.1) / 0.01)))).float().to(dev))
q = P.DenseSdfQuery(meta, tabs[0], tabs[1], dec, axes)
for _ in range(2):
    vol = q.run(0, q.ny)
torch.cuda.synchronize()
ex = meshmod.MeshExtractor(axes)
v, f = ex.run(vol.view(q.ny, q.nx, q.nz))
torch.cuda.synchronize()
print("dense query", vol.numel(), "points; mesh", v.shape[0], "vertices", f.shape[0], "faces")
del vol, v, f
cam = cfg.cam
col, dep, c2w = wl.cur_frame
step = P.RenderImageStep(meta, tabs[0], tabs[1], dec, beta, n_stratified=cfg.n_stratified, n_importance=cfg.n_importance, truncation=cfg.truncation,
                         H=cam.H, W=cam.W, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy)
C, S = step.chunk, step.S
n = cam.H * cam.W
out = step.alloc_outputs(n)
for c0 in range(0, n, C):
    m = min(C, n - c0)
    step.run(c2w, dep, torch.rand(m, S, device=dev), torch.rand(m, cfg.n_stratified, device=dev), torch.rand(m, cfg.n_importance, device=dev),
             pixel_begin=c0, pixel_end=c0 + m, out={k: t[c0:c0 + m] for k, t in out.items()})
torch.cuda.synchronize()
print("render_img", n, "rays; mean depth", float(out["depth"].mean()))
e = cfg.ignore_edge
trk = P.TrackingStep(meta, tabs[0], tabs[1], dec, beta, n_stratified=cfg.n_stratified, n_importance=cfg.n_importance, truncation=cfg.truncation,
                     H=cam.H, W=cam.W, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, ignore_edge_h=e, ignore_edge_w=e, n_rays=cfg.track_pixels)
pose = wlmod._matrix_to_cam_pose(c2w[None]).contiguous()
for _ in range(2):
    trk.run(pose, dep, col, torch.randint((cam.H - 2 * e) * (cam.W - 2 * e), (cfg.track_pixels,), device=dev), torch.rand(cfg.track_pixels, trk.S, device=dev))
torch.cuda.synchronize()
print("tracking loss", float(trk.loss))
