import importlib, sys, dataclasses, json
sys.path.insert(0, "/root/repo")
import torch
P = importlib.import_module("uni-slam_b200"); wlmod = importlib.import_module("uni-slam_b200.workload"); syn = P.synthetic
dev = "cuda:0"
for l2s, l2c in ((16, 19), (8, 8), (12, 12)):
    cfg = dataclasses.replace(syn.REPLICA_ROOM0, log2_hash_sdf=l2s, log2_hash_color=l2c)
    wl = wlmod.build_mapping_workload(cfg, dev, seed=1, scale_hw=0.5)
    meta, tabs, dec, beta = wlmod.init_field_tensors(cfg, wl.bound, wl.per_level_scale, dev)
    step = P.MappingStep(meta, tabs[0], tabs[1], dec, beta, n_stratified=32, n_importance=8, truncation=0.06, max_rays=wl.n_rays, max_frames=wl.K)
    cam_poses = wl.cam_poses.clone()
    step.profile = True
    for it in range(12):
        d = wl.draw()
        if it == 2: step.events = {}
        step.run(wl.batches(d[0], d[1]), d[2], d[3], d[4], cam_poses=cam_poses, c2w_fixed=wl.c2ws[0])
    torch.cuda.synchronize()
    k = step.kernel_ms()
    print(json.dumps({"log2": (l2s, l2c), "fwd_us": round(k["usl_field_fwd"]*1e3,1), "bwd_us": round(k["usl_field_bwd"]*1e3,1)}))
