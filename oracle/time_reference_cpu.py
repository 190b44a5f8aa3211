"""oracle/time_reference_cpu.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

How fast is the UNMODIFIED reference on the host CPU at the bench workload's size, and how does the oracle port that
`bench.py --impl reference` times compare with it?  bench.py may not read /root/reference (it does not exist on the GPU
box), so its CPU arm is the port (`cpu_baseline.kind = "port"`); this script, run in the build container only, times
`Mapper.optimize_mapping` itself (src/Mapper.py:275-459, through oracle/shims: tinycudann -> oracle/grid_ref.py, exactly
as oracle/gen_golden.py drives it) on a 22-frame 1200x680 window = 5982 rays x 40 samples with joint pose optimisation,
and prints one JSON line.  Per-iteration time = (t(1 + n iterations) - t(1 iteration)) / n, so the per-call work
(keyframe stacking, optimiser construction) is excluded -- the port's figure excludes it too.

    python -m oracle.time_reference_cpu [n_iterations]
"""
import json
import os
import sys
import time
import types

import torch

from oracle import gen_golden as G


def build_mapper(case):
    from src.Mapper import Mapper
    from src.utils.Renderer import Renderer
    from src.common import get_camera_rays
    cfg = G._load_cfg(case)
    cfg["mapping"]["pixels"] = case["pixels"]
    bound, grids, dec = G._build_world(cfg, 0)
    H, W = case["H"], case["W"]
    cam = cfg["cam"]
    n_kf = case["n_kf"]
    frames, dirs = G._frames(cfg, case, n_kf + 1, seed=7)
    fake = types.SimpleNamespace(bound=bound, device="cpu", H=H, W=W, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"])
    m = object.__new__(Mapper)
    m.cfg = cfg; m.device = "cpu"; m.truncation = cfg["model"]["truncation"]; m.bound = bound
    m.renderer = Renderer(cfg, fake); m.decoders = dec
    m.hash_grids_xyz = [grids[0]]; m.c_hash_grids_xyz = [grids[1]]
    m.H, m.W, m.fx, m.fy, m.cx, m.cy = H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    for k in ("w_sdf_fs", "w_sdf_center", "w_sdf_tail", "w_depth", "w_color"):
        setattr(m, k, cfg["mapping"][k])
    m.mapping_pixels = cfg["mapping"]["pixels"]; m.mapping_window_size = cfg["mapping"]["mapping_window_size"]
    m.keyframe_selection_method = cfg["mapping"]["keyframe_selection_method"]
    m.m_mask_mode = cfg["m_mask_mode"]; m.no_vis_on_first_frame = True
    m.joint_opt_cam_lr = cfg["mapping"]["joint_opt_cam_lr"]; m.LC = cfg["mapping"]["LC"]
    m.LC_cnt = torch.zeros(1).int(); m.tracking_back = torch.tensor([0])
    m.activated_mapping_mode = cfg["tracking"].get("activated_mapping_mode", False)
    m.visualizer = types.SimpleNamespace(save_imgs=lambda *a, **k: None)
    rays_d_cam = get_camera_rays(H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    torch.manual_seed(11)
    m.keyframe_dict, m.keyframe_list = [], []
    est = torch.zeros(4 * (n_kf + 1) + 1, 4, 4)
    for k in range(n_kf):
        col, dep, c2w = frames[k]
        ind = torch.randperm(H * W)[: int(H * W * 0.1)]
        noisy = c2w.clone(); noisy[:3, 3] += 0.01 * torch.randn(3)
        est[4 * k] = noisy
        m.keyframe_list.append(4 * k)
        m.keyframe_dict.append({"gt_c2w": c2w, "idx": 4 * k, "color": col.reshape(-1, 3)[ind], "depth": dep.reshape(-1)[ind],
                                "est_c2w": noisy.clone(), "rays_d": rays_d_cam.reshape(-1, 3)[ind]})
    m.estimate_c2w_list = est
    m.joint_opt = (len(m.keyframe_list) > 4) and cfg["mapping"]["joint_opt"]
    col, dep, c2w = frames[n_kf]
    cur_c2w = c2w.clone(); cur_c2w[:3, 3] += 0.01
    return m, cfg, (4 * n_kf, col, dep, c2w, cur_c2w, rays_d_cam)


def main():
    G._setup_paths()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    case = dict(yaml="configs/Replica/room0.yaml", H=680, W=1200, s=1.0, n_kf=21, pixels=4000, lr_factor=1)
    m, cfg, (idx, col, dep, c2w, cur_c2w, dirs) = build_mapper(case)
    S = cfg["rendering"]["n_stratified"] + cfg["rendering"]["n_importance"]

    def run(iters):
        torch.manual_seed(2)
        t0 = time.perf_counter()
        m.optimize_mapping(iters, case["lr_factor"], idx, col, dep, c2w, m.keyframe_dict, m.keyframe_list, cur_c2w.clone(), dirs)
        return time.perf_counter() - t0
    run(1)                                                   # warm-up (allocator, thread pool)
    t1 = run(1)
    tn = run(1 + n)
    per_iter = (tn - t1) / n
    K = len(m.keyframe_list) + 1
    rays = K * (case["pixels"] // K) + (10 * 200 if len(m.keyframe_list) > 20 else 0)      # Mapper.py:379-393
    print(json.dumps({"impl": "unmodified reference (src/Mapper.py optimize_mapping through oracle/shims)", "host_threads": threads,
                      "rays": rays, "samples_per_ray": S, "frames": K, "joint_opt": bool(m.joint_opt), "iterations_timed": n,
                      "seconds_per_iteration": per_iter, "ray_samples_per_s": rays * S / per_iter,
                      "seconds_first_call_1_iteration": t1}))


if __name__ == "__main__":
    main()
