"""Run the minimal Tracker+Mapper loop (BASELINE config 2) and print trajectory error and throughput."""
import argparse, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
P = importlib.import_module("uni-slam_b200")
slam = importlib.import_module("uni-slam_b200.slam")
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="replica_room0"); ap.add_argument("--frames", type=int, default=40)
ap.add_argument("--scale", type=float, default=0.5); ap.add_argument("--stride", type=int, default=1); ap.add_argument("--noise", type=float, default=0.0)
a = ap.parse_args()
r = slam.run_slam(P.synthetic.CONFIGS[a.config], n_frames=a.frames, scale_hw=a.scale, frame_stride=a.stride, prior_noise_m=a.noise)
step = (r.gt_c2w[1:, :3, 3] - r.gt_c2w[:-1, :3, 3]).norm(dim=-1)
print(json.dumps({"config": a.config, "frames": a.frames, "scale_hw": a.scale, "ate_rmse_m": r.ate_rmse, "prior_rmse_m": r.ate_rmse_no_tracking,
                  "mean_motion_per_frame_m": float(step.mean()), "path_length_m": float(step.sum()), "frames_per_s": r.frames_per_s,
                  "tracking_iters": r.tracking_iters, "mapping_iters": r.mapping_iters, "mapping_samples": r.mapping_samples,
                  "seconds": r.seconds, "loss_first_map": r.loss_first_map, "loss_last_map": r.loss_last_map}))
