"""Minimal Tracker + Mapper loop on a synthetic RGB-D sequence, built from the fused step drivers.

This is BASELINE.json config 2 ("full Tracker+Mapper loop over a synthetic sequence on 1 B200") as a *driver*:
it reproduces the data flow of src/Tracker.py:271-372 and src/Mapper.py:461-575 -- constant-speed pose
initialisation, N tracking iterations with Adam(T, R; betas 0.5/0.999) keeping the best pose, mapping every
`map_every` frames over all keyframes + the current frame (10 % pixel subsets, new Adam per mapped frame,
joint pose optimisation once > 4 keyframes, 200 px x last-10-frames extra batch once > 20 keyframes) -- in one
process.  The reference's control policy around it (two OS processes, loop-closure keyframe selection,
uncertainty-triggered extra iterations, logging, meshing) is out of scope.
It exists to show the hot path does SLAM end to end (trajectory error vs the analytic ground truth) and to
measure frames/s; it is not part of the drop-in boundary.
"""
import math
import time
from dataclasses import dataclass, field
from typing import List

import torch

from . import synthetic as syn
from . import workload as wlmod
from .keyframes import KeyframeStore
from .optim import FusedAdam
from .steps import MappingStep, TrackingStep


@dataclass
class SlamResult:
    est_c2w: torch.Tensor
    gt_c2w: torch.Tensor
    ate_rmse: float
    ate_rmse_no_tracking: float       # error of the constant-velocity prior alone (what tracking must beat)
    frames_per_s: float
    tracking_iters: int
    mapping_iters: int
    mapping_samples: int
    seconds: float
    loss_first_map: float
    loss_last_map: float


def ate_rmse_aligned(est_c2w: torch.Tensor, gt_c2w: torch.Tensor):
    """The reference's ATE figure (src/tools/eval_ate.py:380-445, evaluate_ate): the estimated camera centres are aligned to
    the ground truth by Horn's closed form (rotation + translation, eval_ate.py:202-236: SVD of the centred cross-covariance,
    reflection fixed through the determinants), and absolute_translational_error.rmse = sqrt(mean |residual|^2) is returned
    with the rotation, the translation and the per-frame residual norms.  Host numpy in double, like the reference."""
    import numpy as np
    e = est_c2w[:, :3, 3].detach().to("cpu", torch.float64).numpy().T          # model (3,n): the trajectory that is moved
    g = gt_c2w[:, :3, 3].detach().to("cpu", torch.float64).numpy().T           # data  (3,n)
    em, gm = e.mean(axis=1, keepdims=True), g.mean(axis=1, keepdims=True)
    w = (e - em) @ (g - gm).T                                                   # sum of outer(model column, data column)
    u, _, vh = np.linalg.svd(w.T)
    s = np.eye(3)
    if np.linalg.det(u) * np.linalg.det(vh) < 0:
        s[2, 2] = -1.0
    rot = u @ s @ vh
    trans = gm - rot @ em
    err = rot @ e + trans - g
    te = np.sqrt((err * err).sum(axis=0))
    return float(np.sqrt(te @ te / len(te))), rot, trans[:, 0], te


def run_slam(cfg: syn.SceneCfg = syn.REPLICA_ROOM0, n_frames: int = 40, device="cuda:0", scale_hw: float = 0.5,
             frame_stride: int = 1, track_iters: int = None, map_iters: int = None, map_iters_first: int = 10,
             seed: int = 0, prior_noise_m: float = 0.0, verbose: bool = False, pregenerate: bool = False,
             graphs: bool = False, graph_mapping: bool = False) -> SlamResult:
    """graphs=True: every tracking iteration after the first is replayed from ONE CUDA graph (RNG draws, the step's launches and
    the fused Adam step); inputs live in static buffers (frame images, draw tensors, pose parameters) and the optimisers are
    reset in place per frame instead of being rebuilt.  graph_mapping=True also captures one graph per mapping window size K
    -- worth it only when window sizes recur.  Same arithmetic as the eager loop.
    Measured on the 200-frame Replica sequence (B200): eager 0.88-1.08 s, tracking graph 1.18 s, all graphs 1.02 s -- the
    eager loop is already GPU-limited (745 mapping iterations x 0.55 ms + 1592 tracking iterations x 0.11 ms = 0.59 s of
    kernels; the launches are asynchronous and stay ahead), so graphs buy nothing here and the default is eager."""
    torch.manual_seed(seed)
    seq = syn.SyntheticSequence(cfg, n_frames=max(200, n_frames * frame_stride), device=device, seed=1, scale_hw=scale_hw)
    cam = seq.cam
    H, W = cam.H, cam.W
    P = int(H * W * 0.1)
    bound = syn.load_bound(cfg.bound_yaml)
    res = syn.grid_resolution(bound, cfg.voxel)
    pls = float(2.0 ** (math.log2(res / 16) / 15))
    meta, tabs, dec, beta = wlmod.init_field_tensors(cfg, bound, pls, device, seed=seed)
    S = cfg.n_stratified + cfg.n_importance
    track_iters = track_iters or cfg.track_iters
    map_iters = map_iters or cfg.map_iters
    edge = max(int(cfg.ignore_edge * scale_hw), 2)
    max_kf = n_frames // cfg.map_every + 3
    max_rays = cfg.map_pixels + 2000 + 64
    mstep = MappingStep(meta, tabs[0], tabs[1], dec, beta, n_stratified=cfg.n_stratified, n_importance=cfg.n_importance,
                        truncation=cfg.truncation, max_rays=max_rays, max_frames=max_kf)
    tstep = TrackingStep(meta, tabs[0], tabs[1], dec, beta, n_stratified=cfg.n_stratified, n_importance=cfg.n_importance,
                         truncation=cfg.truncation, H=H, W=W, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy,
                         ignore_edge_h=edge, ignore_edge_w=edge, n_rays=cfg.track_pixels)
    params = [tabs[0], tabs[1], beta] + dec
    grads = [mstep.fs.g_sdf_table, mstep.fs.g_rgb_table, mstep.fs.g_beta] + mstep.fs.g_dec
    for p_, g_ in zip(params, grads):
        p_.requires_grad_(True); p_.grad = g_
    store = KeyframeStore(max_kf, H, W, device)         # Mapper.py:528-541: 10 % pixel subsets, device resident
    est = torch.zeros((n_frames, 4, 4), device=device); gt = torch.zeros((n_frames, 4, 4), device=device)
    prior = torch.zeros((n_frames, 4, 4), device=device)
    n_track = n_map = map_samples = 0
    loss_first = loss_last = float("nan")
    to_pose = wlmod._matrix_to_cam_pose
    npx_win = (H - 2 * edge) * (W - 2 * edge)
    # pregenerate: render the synthetic RGB-D frames before the clock starts (they stand in for the dataset on disk, which the
    # reference's loader threads read ahead of the tracker); otherwise frame synthesis is part of the measured loop
    frames = [seq.frame(k * frame_stride) for k in range(n_frames)] if pregenerate else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if graphs:
        return _run_graphed(locals())
    for k in range(n_frames):
        col, dep, c2w_gt = frames[k] if frames is not None else seq.frame(k * frame_stride)
        gt[k] = c2w_gt
        # ---------------- tracking (Tracker.py:306-368) ----------------
        if k == 0:
            est[0] = c2w_gt; prior[0] = c2w_gt
        else:
            if k >= 2:
                pp = to_pose(torch.stack([est[k - 2], est[k - 1]]))
                pose0 = 2 * pp[1:] - pp[0:1]                                   # constant-speed assumption (Tracker.py:315-319)
            else:
                pose0 = to_pose(est[k - 1][None])
            if prior_noise_m > 0:                                              # perturbed initial guess: the tracker has to pull it back
                pose0 = pose0.clone(); pose0[:, 4:] += prior_noise_m * torch.randn(3, device=device)
            prior[k] = _pose_to_c2w(pose0)
            cam_pose = pose0.clone().contiguous()                              # cat([R, T]) (Tracker.py:333): R, T are its two halves
            T_ = cam_pose[:, 4:].requires_grad_(True); R_ = cam_pose[:, :4].requires_grad_(True)
            T_.grad = tstep.d_pose[:, 4:]; R_.grad = tstep.d_pose[:, :4]
            opt = FusedAdam([{"params": [T_], "lr": cfg.lr_T, "betas": (0.5, 0.999)}, {"params": [R_], "lr": cfg.lr_R, "betas": (0.5, 0.999)}])
            best_loss = torch.full((1,), float("inf"), device=device); best_pose = pose0.clone()
            for _ in range(track_iters):
                idx = torch.randint(npx_win, (cfg.track_pixels,), device=device)
                t_rand = torch.rand((cfg.track_pixels, S), device=device)
                tstep.run(cam_pose, dep, col, idx, t_rand, best_loss, best_pose)   # keeps the best pose on the device (Tracker.py:346-348)
                opt.step()
                n_track += 1
            est[k] = _pose_to_c2w(best_pose)
        # ---------------- mapping (Mapper.py:485-541) ----------------
        if k % cfg.map_every == 0:
            n_kf = len(store)
            K = n_kf + 1
            store.stage_current(col, dep, seq.dirs, est[k], c2w_gt, indices=torch.randperm(H * W, device=device)[:P])
            kf_c2w = store.est_c2w
            joint = n_kf > 4                                                  # Mapper.py:519
            first = k == 0
            lr_f = 5.0 if first else 1.0                                       # lr_first_factor / lr_factor
            groups = [{"params": dec + [beta], "lr": 1e-3 * lr_f}, {"params": [tabs[0]], "lr": cfg.hash_lr * lr_f},
                      {"params": [tabs[1]], "lr": cfg.hash_lr * lr_f}]
            cam_poses = None
            if joint:
                cam_poses = to_pose(kf_c2w[1:K]).contiguous().requires_grad_(True)
                cam_poses.grad = mstep.d_pose[:K - 1]
                groups.append({"params": [cam_poses], "lr": 1e-3})
            opt = FusedAdam(groups)                                            # new Adam per mapped frame (Mapper.py:364)
            n_main = cfg.map_pixels // K
            n_rec = 200 if n_kf > 20 else 0
            R = K * n_main + 10 * n_rec
            iters = map_iters_first if first else map_iters
            for it in range(iters):
                idx_main = torch.randint(P, (K * n_main,), device=device)
                idx_rec = torch.randint(P, (10 * n_rec,), device=device) if n_rec else None
                batches = store.mapping_batches(idx_main, n_main, idx_rec, n_rec)   # views into the store, no stacking
                loss = mstep.run(batches, torch.rand((R, S), device=device), torch.rand((R, cfg.n_stratified), device=device),
                                 torch.rand((R, cfg.n_importance), device=device),
                                 cam_poses=cam_poses.detach() if joint else None, c2w_fixed=kf_c2w[0] if joint else None)
                opt.step()
                n_map += 1; map_samples += R * S
                if first and it == 0:
                    loss_first = float(loss)
            loss_last = float(loss)
            if joint:                                                          # Mapper.py:447-457
                est[k] = store.write_back_poses(_pose_to_c2w(cam_poses.detach()))
            store.promote_staged(k)                                            # keyframe_every == map_every in every config
            if verbose:
                print(f"frame {k}: mapped K={K} loss={loss_last:.4f}")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ate = float((est[:, :3, 3] - gt[:, :3, 3]).pow(2).sum(-1).mean().sqrt())
    ate_prior = float((prior[:, :3, 3] - gt[:, :3, 3]).pow(2).sum(-1).mean().sqrt())
    return SlamResult(est, gt, ate, ate_prior, n_frames / dt, n_track, n_map, map_samples, dt, loss_first, loss_last)


def _pose_to_c2w(pose: torch.Tensor) -> torch.Tensor:
    from . import ops
    out = ops.pose_to_matrix(pose.contiguous())
    return out[0] if pose.shape[0] == 1 else out


def _capture(fn):
    """One eager call of fn has just run; capture a second one (capture does not execute) and return the replay.
    (capture_begin / capture_end directly: the torch.cuda.graph context manager runs gc.collect() and empties the allocator
    cache on entry, tens of milliseconds per capture.)"""
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g.capture_begin()
        try:
            fn()
        finally:
            g.capture_end()
    torch.cuda.current_stream().wait_stream(s)
    return g.replay


def _run_graphed(v) -> SlamResult:
    """The loop of run_slam with graph replay (see its docstring).  `v` = run_slam's locals at the start of the timed loop."""
    cfg, n_frames, device, frame_stride, prior_noise_m, verbose = v["cfg"], v["n_frames"], v["device"], v["frame_stride"], v["prior_noise_m"], v["verbose"]
    seq, frames, store, mstep, tstep, tabs, dec, beta = v["seq"], v["frames"], v["store"], v["mstep"], v["tstep"], v["tabs"], v["dec"], v["beta"]
    H, W, P, S, npx_win, to_pose = v["H"], v["W"], v["P"], v["S"], v["npx_win"], v["to_pose"]
    track_iters, map_iters, map_iters_first = v["track_iters"], v["map_iters"], v["map_iters_first"]
    est, gt, prior, t0 = v["est"], v["gt"], v["prior"], v["t0"]
    n_track = n_map = map_samples = 0
    loss_first = loss_last = float("nan")
    # ---- tracking: static inputs + one optimiser, reset per frame ----
    dep_s = torch.empty((H, W), device=device); col_s = torch.empty((H, W, 3), device=device)
    cam_pose = torch.zeros((1, 7), device=device)
    T_ = cam_pose[:, 4:].requires_grad_(True); R_ = cam_pose[:, :4].requires_grad_(True)
    T_.grad = tstep.d_pose[:, 4:]; R_.grad = tstep.d_pose[:, :4]
    opt_t = FusedAdam([{"params": [T_], "lr": cfg.lr_T, "betas": (0.5, 0.999)}, {"params": [R_], "lr": cfg.lr_R, "betas": (0.5, 0.999)}])
    opt_t.enable_graph_step_counter(device)
    t_idx = torch.empty((cfg.track_pixels,), device=device, dtype=torch.int64); t_rand = torch.empty((cfg.track_pixels, S), device=device)
    best_loss = torch.empty((1,), device=device); best_pose = torch.empty((1, 7), device=device)

    def track_it():
        t_idx.random_(0, npx_win); t_rand.uniform_()
        tstep.run(cam_pose, dep_s, col_s, t_idx, t_rand, best_loss, best_pose)
        opt_t.step()
    track_replay = None
    # ---- mapping: one pose parameter block for every window size (rows past K-1 see zero gradients and stay put), two
    # optimisers (first frame: lr x 5, Mapper.py lr_first_factor), one graph per K ----
    poses_all = torch.zeros_like(mstep.d_pose).requires_grad_(True)
    poses_all.grad = mstep.d_pose

    def map_opt(f):
        o = FusedAdam([{"params": dec + [beta], "lr": 1e-3 * f}, {"params": [tabs[0]], "lr": cfg.hash_lr * f},
                       {"params": [tabs[1]], "lr": cfg.hash_lr * f}, {"params": [poses_all], "lr": 1e-3}])
        o.enable_graph_step_counter(device)
        return o
    opt_first, opt_rest = map_opt(5.0), map_opt(1.0)
    map_replay = {}
    graph_mapping = v["graph_mapping"]
    for k in range(n_frames):
        col, dep, c2w_gt = frames[k] if frames is not None else seq.frame(k * frame_stride)
        gt[k] = c2w_gt
        if k == 0:
            est[0] = c2w_gt; prior[0] = c2w_gt
        else:
            if k >= 2:
                pp = to_pose(torch.stack([est[k - 2], est[k - 1]]))
                pose0 = 2 * pp[1:] - pp[0:1]
            else:
                pose0 = to_pose(est[k - 1][None])
            if prior_noise_m > 0:
                pose0 = pose0.clone(); pose0[:, 4:] += prior_noise_m * torch.randn(3, device=device)
            prior[k] = _pose_to_c2w(pose0)
            with torch.no_grad():
                cam_pose.copy_(pose0); best_pose.copy_(pose0); best_loss.fill_(float("inf"))
                dep_s.copy_(dep); col_s.copy_(col)
            opt_t.reset_state()
            start = 0
            if track_replay is None:
                track_it(); start = 1
                track_replay = _capture(track_it)
            for _ in range(start, track_iters):
                track_replay()
            n_track += track_iters
            est[k] = _pose_to_c2w(best_pose)
        if k % cfg.map_every == 0:
            n_kf = len(store)
            K = n_kf + 1
            store.stage_current(col, dep, seq.dirs, est[k], c2w_gt, indices=torch.randperm(H * W, device=device)[:P])
            kf_c2w = store.est_c2w
            joint = n_kf > 4
            first = k == 0
            opt = opt_first if first else opt_rest
            opt.reset_state()
            n_main = cfg.map_pixels // K
            n_rec = 200 if n_kf > 20 else 0
            R = K * n_main + 10 * n_rec
            iters = map_iters_first if first else map_iters
            if joint:
                with torch.no_grad():
                    poses_all[:K - 1].copy_(to_pose(kf_c2w[1:K]))
            start = 0
            if K not in map_replay:
                m_idx = torch.empty((K * n_main,), device=device, dtype=torch.int64)
                m_rec = torch.empty((10 * n_rec,), device=device, dtype=torch.int64) if n_rec else None
                m_tr = torch.empty((R, S), device=device); m_tu = torch.empty((R, cfg.n_stratified), device=device)
                m_up = torch.empty((R, cfg.n_importance), device=device)
                cam_view = poses_all.detach()[:K - 1] if joint else None
                fixed = kf_c2w[0] if joint else None

                def map_it(m_idx=m_idx, m_rec=m_rec, m_tr=m_tr, m_tu=m_tu, m_up=m_up, cam_view=cam_view, fixed=fixed, n_main=n_main, n_rec=n_rec, opt=opt):
                    m_idx.random_(0, P); m_tr.uniform_(); m_tu.uniform_(); m_up.uniform_()
                    if m_rec is not None:
                        m_rec.random_(0, P)
                    mstep.run(store.mapping_batches(m_idx, n_main, m_rec, n_rec), m_tr, m_tu, m_up, cam_poses=cam_view, c2w_fixed=fixed)
                    opt.step()
                map_it(); start = 1
                if first:
                    loss_first = float(mstep.loss)
                # a window size that comes back (bounded windows) replays a graph; one that occurs once (this driver maps over
                # ALL keyframes, so K grows by one per mapped frame) runs eagerly: instantiating a graph costs more than the
                # 14 launches-worth of host time it would save
                map_replay[K] = _capture(map_it) if graph_mapping else map_it
            for _ in range(start, iters):
                map_replay[K]()
            n_map += iters; map_samples += iters * R * S
            if joint:
                est[k] = store.write_back_poses(_pose_to_c2w(poses_all.detach()[:K - 1]))
            store.promote_staged(k)
            if verbose:
                print(f"frame {k}: mapped K={K} loss={float(mstep.loss):.4f}")
    loss_last = float(mstep.loss)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ate = float((est[:, :3, 3] - gt[:, :3, 3]).pow(2).sum(-1).mean().sqrt())
    ate_prior = float((prior[:, :3, 3] - gt[:, :3, 3]).pow(2).sum(-1).mean().sqrt())
    return SlamResult(est, gt, ate, ate_prior, n_frames / dt, n_track, n_map, map_samples, dt, loss_first, loss_last)
