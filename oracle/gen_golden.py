"""oracle/gen_golden.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Generates tests/golden/*.npz by running the UNMODIFIED reference Python code
(/root/reference/src/{Mapper,Tracker,common}.py, src/utils/Renderer.py,
src/networks/decoders.py) on CPU through oracle/shims, recording every RNG draw, the inputs
and outputs of the hot-path seams (get_samples / get_samples_all / render_batch_ray), the loss
and the gradients Adam sees.  Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden

The fixtures pin oracle/path_ref.py (tests/test_oracle_golden.py) and are compared directly with
the CUDA path in the -m gpu tests.  Grid tables and decoder weights are filled with the
platform-independent integer-hash generator grid_ref.lcg_params, so fixtures carry no tables.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.environ.get("USL_GOLDEN_OUT", os.path.join(REPO, "tests", "golden"))    # override: tests/test_golden_reproduce.py

CASES = {
    # name: (yaml, decoder variant implied by yaml, H, W, intrinsics scale, n_keyframes, joint_opt)
    "map_replica_k1": dict(yaml="configs/Replica/room0.yaml", H=60, W=80, s=1 / 15.0, n_kf=0, pixels=240, lr_factor=5),
    "map_replica_k7": dict(yaml="configs/Replica/room0.yaml", H=60, W=80, s=1 / 15.0, n_kf=6, pixels=280, lr_factor=1),
    "map_scannet_k23": dict(yaml="configs/ScanNet/scene0000.yaml", H=46, W=62, s=0.1, n_kf=22, pixels=230, lr_factor=1, threads=1),
    "track_replica": dict(yaml="configs/Replica/room0.yaml", H=60, W=80, s=1 / 15.0, pixels=200, edge=6),
    "track_scannet": dict(yaml="configs/ScanNet/scene0000.yaml", H=46, W=62, s=0.1, pixels=200, edge=5),
    # m_mask_mode / t_mask_mode "no_mask" (Mapper.py:432-440, Tracker.py:230-238)
    "map_replica_nomask": dict(yaml="configs/Replica/room0.yaml", H=60, W=80, s=1 / 15.0, n_kf=0, pixels=160, lr_factor=5, mask_mode="no_mask"),
    "track_scannet_nomask": dict(yaml="configs/ScanNet/scene0000.yaml", H=46, W=62, s=0.1, pixels=120, edge=5, mask_mode="no_mask"),
    # keyframe bookkeeping: full (tiny) frames + every randperm draw, so the device-resident KeyframeStore can be checked
    # against the tensors optimize_mapping stacks from its list of dicts (Mapper.py:315-351)
    "map_replica_kfstore": dict(yaml="configs/Replica/room0.yaml", H=30, W=40, s=1 / 30.0, n_kf=3, pixels=120, lr_factor=1, save_frames=True),
    # Mesher.get_grid_uniform + eval_points (Mesher.py:134-195): the dense SDF query seam, coarse grids (ragged counts)
    "mesh_replica": dict(yaml="configs/Replica/room0.yaml", H=30, W=40, s=1 / 30.0, resolution=0.37),
    "mesh_scannet": dict(yaml="configs/ScanNet/scene0000.yaml", H=23, W=31, s=0.05, resolution=0.53),
    # Mapper.keyframe_selection_LC (Mapper.py:177-236): the co-visibility measure between the current frame and every keyframe
    "kf_covis_replica": dict(yaml="configs/Replica/room0.yaml", H=60, W=80, s=1 / 15.0, n_kf=9),
    # Renderer.render_img (Renderer.py:160-223): whole frame in ray_batch_size chunks, last chunk ragged
    "img_replica": dict(yaml="configs/Replica/room0.yaml", H=30, W=40, s=1 / 30.0, ray_batch=500),
    "img_scannet": dict(yaml="configs/ScanNet/scene0000.yaml", H=23, W=31, s=0.05, ray_batch=300),
    # src/tools/cull_mesh.py: cull_mesh (frustum + occlusion test against the frames, both eval_rec settings) and
    # cull_out_bound_mesh (convex bound) on a small marching-cubes mesh of the analytic room
    "cull_replica": dict(yaml="configs/Replica/room0.yaml", H=30, W=40, s=1 / 30.0, n_frames=7, voxel=0.31),
    # src/tools/eval_recon.py:235-307 eval_rendering: render_img on every 5th frame, PSNR over the pixels with depth, depth L1
    "evalr_replica": dict(yaml="configs/Replica/room0.yaml", H=30, W=40, s=1 / 30.0, n_img=11, ray_batch=700),
}


def _setup_paths():
    sys.path[:0] = [os.path.join(HERE, "shims"), REF, REPO]
    os.chdir(REF)   # the yaml files use repo-relative inherit_from paths (src/config.py:38-43)


class Recorder:
    def __init__(self):
        self.draws = []       # (kind, tensor)
        self.samples = []     # get_samples / get_samples_all calls
        self.renders = []
        self.losses = []
        self.steps = []       # list of list-of-grads per optimizer.step()

    def install(self):
        import src.Mapper as M
        import src.Tracker as T
        rec = self
        self._orig = dict(rand=torch.rand, randint=torch.randint, randperm=torch.randperm,
                          backward=torch.Tensor.backward, step=torch.optim.Adam.step,
                          gsa=M.get_samples_all, gs=T.get_samples, searchsorted=torch.searchsorted)
        self.searches = []    # outputs of torch.searchsorted (sample_pdf, common.py:70): the "sample indices" of the no-depth branch

        def searchsorted(*a, **k):
            t = rec._orig["searchsorted"](*a, **k); rec.searches.append(t.clone()); return t
        torch.searchsorted = searchsorted

        def rand(*a, **k):
            t = rec._orig["rand"](*a, **k); rec.draws.append(("rand", t.clone())); return t

        def randint(*a, **k):
            t = rec._orig["randint"](*a, **k); rec.draws.append(("randint", t.clone())); return t

        def randperm(*a, **k):
            t = rec._orig["randperm"](*a, **k); rec.draws.append(("randperm", t.clone())); return t

        def backward(self_t, *a, **k):
            rec.losses.append(self_t.detach().clone()); return rec._orig["backward"](self_t, *a, **k)

        def step(opt, *a, **k):
            rec.steps.append([[None if p.grad is None else p.grad.detach().clone() for p in g["params"]]
                              for g in opt.param_groups])
            rec.params_at_step = [[p.detach().clone() for p in g["params"]] for g in opt.param_groups]
            return rec._orig["step"](opt, *a, **k)

        def gsa(*a, **k):
            out = rec._orig["gsa"](*a, **k)
            rec.samples.append(dict(kind="all", args=[x.detach().clone() if torch.is_tensor(x) else x for x in a],
                                    out=[o.detach().clone() for o in out]))
            return out

        def gs(*a, **k):
            out = rec._orig["gs"](*a, **k)
            rec.samples.append(dict(kind="win", args=[x.detach().clone() if torch.is_tensor(x) else x for x in a],
                                    out=[o.detach().clone() for o in out]))
            return out

        torch.rand, torch.randint, torch.randperm = rand, randint, randperm
        torch.Tensor.backward = backward
        torch.optim.Adam.step = step
        M.get_samples_all = gsa
        T.get_samples = gs
        self._m_gs = getattr(M, "get_samples", None)      # keyframe_selection_LC samples through Mapper's own import of it
        if self._m_gs is not None:
            def mgs(*a, **k):
                out = rec._m_gs(*a, **k)
                rec.samples.append(dict(kind="win", args=[x.detach().clone() if torch.is_tensor(x) else x for x in a],
                                        out=[o.detach().clone() for o in out]))
                return out
            M.get_samples = mgs

    def uninstall(self):
        import src.Mapper as M
        import src.Tracker as T
        torch.rand, torch.randint, torch.randperm = self._orig["rand"], self._orig["randint"], self._orig["randperm"]
        torch.searchsorted = self._orig["searchsorted"]
        torch.Tensor.backward = self._orig["backward"]
        torch.optim.Adam.step = self._orig["step"]
        M.get_samples_all = self._orig["gsa"]
        T.get_samples = self._orig["gs"]
        if self._m_gs is not None:
            M.get_samples = self._m_gs


def _load_cfg(case):
    from src import config
    cfg = config.load_config(case["yaml"], "configs/UNISLAM.yaml")
    cfg["device"] = "cpu"; cfg["keyframe_device"] = "cpu"
    s = case["s"]
    cam = cfg["cam"]
    # the reference applies crop_edge in UNISLAM.update_cam; here we directly give the post-crop camera, scaled
    edge = cam["crop_edge"]
    cam.update(H=case["H"], W=case["W"], fx=cam["fx"] * s, fy=cam["fy"] * s, cx=(cam["cx"] - edge) * s,
               cy=(cam["cy"] - edge) * s, crop_edge=0)
    return cfg


def _build_world(cfg, seed_salt):
    """Grids (via the reference's own get_encoder arithmetic), decoders (reference class), bound."""
    import tinycudann as tcnn
    from src.networks.decoders import Decoders
    from oracle import grid_ref, path_ref
    bound = path_ref.load_bound(cfg["mapping"]["bound"], cfg["scale"], cfg["planes_res"]["bound_dividable"])
    dim_max = (bound[:, 1] - bound[:, 0]).max()
    grids = []
    for key_hash, key_vox, salt in (("hash_size_sdf", "voxel_sdf", 1), ("hash_size_color", "voxel_color", 2)):
        res = int(dim_max / cfg["grid"][key_vox])                                   # UNISLAM.py:196-199
        pls = np.exp2(np.log2(res / 16) / (16 - 1))                                  # UNISLAM.py:241
        enc = tcnn.Encoding(n_input_dims=3, encoding_config={
            "otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2,
            "log2_hashmap_size": cfg["grid"][key_hash], "base_resolution": 16, "per_level_scale": pls}, dtype=torch.float)
        with torch.no_grad():
            enc.params.copy_(torch.from_numpy(grid_ref.lcg_params(enc.params.numel(), 0.05, salt + seed_salt)))
        grids.append(enc)
    dec = Decoders(cfg, c_dim=cfg["model"]["c_dim"], truncation=cfg["model"]["truncation"],
                   learnable_beta=cfg["rendering"]["learnable_beta"])
    dec.bound = bound
    with torch.no_grad():
        for k, (name, p) in enumerate(sorted(dec.named_parameters())):
            if name == "beta":
                continue
            fan_in = p.shape[-1] if p.dim() > 1 else 16
            sc = 1.0 / np.sqrt(fan_in) if p.numel() != 768 else 0.35
            p.copy_(torch.from_numpy(grid_ref.lcg_params(p.numel(), sc, 100 + k + seed_salt)).reshape(p.shape))
    return bound, grids, dec


def _frames(cfg, case, n, seed):
    """Tiny synthetic frames from the oracle's own frozen scene (oracle/scene.py): nothing in the product package
    can change what this generator writes."""
    from oracle import scene as syn
    cam = syn.CameraCfg(case["H"], case["W"], cfg["cam"]["fx"], cfg["cam"]["fy"], cfg["cam"]["cx"], cfg["cam"]["cy"])
    room = syn.AnalyticRoom(cfg["mapping"]["bound"])
    poses = syn.trajectory(room, 200)
    dirs = syn.camera_dirs(cam)
    g = torch.Generator().manual_seed(seed)
    out = []
    for k in range(n):
        c2w = poses[(k * 4) % 200]
        d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
        t, p = room.trace(c2w[:3, 3].expand(d.shape), d)
        col = torch.where((t > 0)[..., None], room.albedo(p), torch.zeros(3))
        hole = torch.rand(t.shape, generator=g) < 0.05
        t = torch.where(hole, torch.zeros_like(t), t)
        out.append((col.float().contiguous(), t.float().contiguous(), c2w.clone()))
    return out, dirs


def _save(name, out):
    """Deterministic container bytes are not needed (the reproduce test compares arrays), but keep key order stable."""
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: out[k] for k in sorted(out)})


def _np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def _grad_summary(prefix, g, spec, out):
    g64 = g.double().reshape(-1, 2)
    lv_sum, lv_norm = [], []
    for lv in spec.levels:
        seg = g64[lv.offset: lv.offset + lv.size]
        lv_sum.append(seg.sum().item()); lv_norm.append(seg.norm().item())
    nz = torch.nonzero(g.reshape(-1)).reshape(-1)
    pick = nz[torch.linspace(0, nz.numel() - 1, min(4096, nz.numel())).long()] if nz.numel() else nz
    out[prefix + "_level_sum"] = np.array(lv_sum); out[prefix + "_level_norm"] = np.array(lv_norm)
    out[prefix + "_nnz"] = np.array(nz.numel()); out[prefix + "_idx"] = _np(pick); out[prefix + "_val"] = _np(g.reshape(-1)[pick])


def gen_mapping(name, case):
    from src.Mapper import Mapper
    from src.utils.Renderer import Renderer
    from src.common import get_camera_rays
    cfg = _load_cfg(case)
    cfg["m_mask_mode"] = case.get("mask_mode", cfg["m_mask_mode"])
    cfg["mapping"]["pixels"] = case["pixels"]
    bound, grids, dec = _build_world(cfg, 0)
    H, W = case["H"], case["W"]
    cam = cfg["cam"]
    n_kf = case["n_kf"]
    frames, dirs = _frames(cfg, case, n_kf + 1, seed=7)
    fake = types.SimpleNamespace(bound=bound, device="cpu", H=H, W=W, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"])
    renderer = Renderer(cfg, fake)
    m = object.__new__(Mapper)
    m.cfg = cfg; m.device = "cpu"; m.truncation = cfg["model"]["truncation"]; m.bound = bound
    m.renderer = renderer; m.decoders = dec
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
    assert torch.equal(rays_d_cam, dirs)
    # keyframe store exactly as Mapper.run builds it (Mapper.py:528-541)
    torch.manual_seed(11)
    m.keyframe_dict, m.keyframe_list = [], []
    est = torch.zeros(4 * (n_kf + 1) + 1, 4, 4)
    kf_inds = []
    for k in range(n_kf):
        col, dep, c2w = frames[k]
        idx = 4 * k
        ind = torch.randperm(H * W)[: int(H * W * 0.1)]
        kf_inds.append(ind)
        noisy = c2w.clone(); noisy[:3, 3] += 0.01 * torch.randn(3)
        est[idx] = noisy
        m.keyframe_list.append(idx)
        m.keyframe_dict.append({"gt_c2w": c2w, "idx": idx, "color": col.reshape(-1, 3)[ind], "depth": dep.reshape(-1)[ind],
                                "est_c2w": noisy.clone(), "rays_d": rays_d_cam.reshape(-1, 3)[ind]})
    m.estimate_c2w_list = est
    m.joint_opt = (len(m.keyframe_list) > 4) and cfg["mapping"]["joint_opt"]           # Mapper.py:519
    col, dep, c2w = frames[n_kf]
    cur_idx = 4 * n_kf
    cur_c2w = c2w.clone(); cur_c2w[:3, 3] += 0.01
    rec = Recorder(); rec.install()
    orig_render = renderer.render_batch_ray

    def render(*a, **k):
        out = orig_render(*a, **k)
        rec.renders.append(dict(rays_d=a[2].detach().clone(), rays_o=a[3].detach().clone(),
                                gt_depth=k["gt_depth"].detach().clone(), out=[o.detach().clone() for o in out]))
        return out
    renderer.render_batch_ray = render
    try:
        torch.manual_seed(2)
        m.optimize_mapping(1, case["lr_factor"], cur_idx, col, dep, c2w, m.keyframe_dict, m.keyframe_list, cur_c2w, rays_d_cam)
    finally:
        rec.uninstall()
    out = {}
    out["meta_H_W_fx_fy_cx_cy"] = np.array([H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"]], dtype=np.float64)
    out["bound"] = _np(bound); out["truncation"] = np.array(m.truncation)
    out["n_stratified"] = np.array(cfg["rendering"]["n_stratified"]); out["n_importance"] = np.array(cfg["rendering"]["n_importance"])
    out["log2_hash"] = np.array([cfg["grid"]["hash_size_sdf"], cfg["grid"]["hash_size_color"]])
    out["per_level_scale"] = np.array([grids[0].spec.per_level_scale, grids[1].spec.per_level_scale])
    out["variant"] = np.array("B" if cfg["grid"]["tcnn_network"] else "A")
    out["joint_opt"] = np.array(int(m.joint_opt)); out["mask_mode"] = np.array(m.m_mask_mode)
    # sampling calls (first = main batch; optional second = the 200px x last-10-frames batch)
    calls = [s for s in rec.samples if s["kind"] == "all"]
    out["n_sample_calls"] = np.array(len(calls))
    randints = [t for k, t in rec.draws if k == "randint"]
    rands = [t for k, t in rec.draws if k == "rand"]
    for ci, s in enumerate(calls):
        a = s["args"]
        out[f"call{ci}_n"] = np.array(a[4]); out[f"call{ci}_c2ws"] = _np(a[11]); out[f"call{ci}_depths"] = _np(a[12])
        out[f"call{ci}_colors"] = _np(a[13]); out[f"call{ci}_rays_d_cam"] = _np(a[15])
        out[f"call{ci}_indices"] = _np(randints[len(randints) - len(calls) + ci])
        for nm, o in zip(("rays_o", "rays_d", "depth", "color"), s["out"]):
            out[f"call{ci}_out_{nm}"] = _np(o)
    r = rec.renders[-1]
    n_valid = int((r["gt_depth"] > 0).sum()); n0 = r["gt_depth"].numel() - n_valid
    tail = rands[-3:] if n0 > 0 else rands[-1:]
    out["t_rand"] = _np(tail[0]); assert tail[0].shape[0] == n_valid
    if n0 > 0:
        out["t_rand_uni"] = _np(tail[1]); out["u_pdf"] = _np(tail[2])
        out["pdf_inds"] = _np(rec.searches[-1]); assert rec.searches[-1].shape == (n0, cfg["rendering"]["n_importance"])
    out["render_rays_o"] = _np(r["rays_o"]); out["render_rays_d"] = _np(r["rays_d"]); out["render_gt_depth"] = _np(r["gt_depth"])
    for nm, o in zip(("term", "pixel_unc", "depth", "rgb", "sdf", "z_vals", "depth_unc"), r["out"]):
        out["ret_" + nm] = _np(o)
    out["loss"] = _np(rec.losses[-1])
    if case.get("save_frames"):
        out["frames_color"] = np.stack([_np(f[0]) for f in frames]); out["frames_depth"] = np.stack([_np(f[1]) for f in frames])
        out["frames_gt_c2w"] = np.stack([_np(f[2]) for f in frames]); out["dirs_cam"] = _np(rays_d_cam)
        out["kf_indices"] = np.stack([_np(i) for i in kf_inds]); out["kf_frame_idx"] = np.array(m.keyframe_list)
        out["kf_est_c2w"] = np.stack([_np(d["est_c2w"]) for d in m.keyframe_dict]); out["cur_c2w"] = _np(cur_c2w)
        perms = [t for k, t in rec.draws if k == "randperm"]
        assert len(perms) == 1                                    # the current frame's subset (Mapper.py:332)
        out["cur_randperm"] = _np(perms[0])
    grads = rec.steps[-1]; pvals = rec.params_at_step
    dec_names = [n for n, _ in dec.named_parameters()]
    for n_, g in zip(dec_names, grads[0]):
        out["grad_dec." + n_] = _np(g)
    _grad_summary("grad_sdf_table", grads[1][0], grids[0].spec, out)
    _grad_summary("grad_rgb_table", grads[2][0], grids[1].spec, out)
    if m.joint_opt:
        out["cam_poses"] = _np(pvals[3][0]); out["grad_cam_poses"] = _np(grads[3][0])
    os.makedirs(OUT, exist_ok=True)
    _save(name, out)
    print(name, "rays", r["gt_depth"].numel(), "holes", n0, "loss", float(rec.losses[-1]), "calls", len(calls))


def gen_tracking(name, case):
    from src.Tracker import Tracker
    from src.utils.Renderer import Renderer
    from src.common import matrix_to_cam_pose
    cfg = _load_cfg(case)
    cfg["t_mask_mode"] = case.get("mask_mode", cfg["t_mask_mode"])
    bound, grids, dec = _build_world(cfg, 50)
    H, W = case["H"], case["W"]
    cam = cfg["cam"]
    frames, dirs = _frames(cfg, case, 2, seed=9)
    fake = types.SimpleNamespace(bound=bound, device="cpu", H=H, W=W, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"])
    renderer = Renderer(cfg, fake)
    t = object.__new__(Tracker)
    t.cfg = cfg; t.device = "cpu"; t.truncation = cfg["model"]["truncation"]; t.bound = bound
    t.renderer = renderer; t.decoders = dec
    for p in t.decoders.parameters():
        p.requires_grad_(False)                                                        # Tracker.py:110-111
    t.hash_grids_xyz = [grids[0]]; t.c_hash_grids_xyz = [grids[1]]
    t.H, t.W, t.fx, t.fy, t.cx, t.cy = H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    for k in ("w_sdf_fs", "w_sdf_center", "w_sdf_tail", "w_depth", "w_color"):
        setattr(t, k, cfg["tracking"][k])
    t.ignore_edge_H = t.ignore_edge_W = case["edge"]
    t.t_mask_mode = cfg["t_mask_mode"]
    col, dep, c2w = frames[1]
    c2w_init = c2w.clone(); c2w_init[:3, 3] += torch.tensor([0.012, -0.008, 0.005])
    cam_pose0 = matrix_to_cam_pose(c2w_init.unsqueeze(0))
    cam_pose0[:, :4] += torch.tensor([0.002, -0.003, 0.001, 0.002])                   # un-normalised quaternion on purpose
    T = torch.nn.Parameter(cam_pose0[:, -3:].clone()); R = torch.nn.Parameter(cam_pose0[:, :4].clone())
    opt = torch.optim.Adam([{"params": [T], "lr": cfg["tracking"]["lr_T"], "betas": (0.5, 0.999)},
                            {"params": [R], "lr": cfg["tracking"]["lr_R"], "betas": (0.5, 0.999)}])   # Tracker.py:324-329
    rec = Recorder(); rec.install()
    orig_render = renderer.render_batch_ray

    def render(*a, **k):
        out = orig_render(*a, **k)
        rec.renders.append(dict(rays_d=a[2].detach().clone(), rays_o=a[3].detach().clone(),
                                gt_depth=k["gt_depth"].detach().clone(), out=[o.detach().clone() for o in out]))
        return out
    renderer.render_batch_ray = render
    try:
        torch.manual_seed(3)
        cam_pose = torch.cat([R, T], -1)
        loss, punc = t.optimize_tracking(cam_pose, col.unsqueeze(0), dep.unsqueeze(0), case["pixels"], opt)
    finally:
        rec.uninstall()
    out = {}
    out["meta_H_W_fx_fy_cx_cy"] = np.array([H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"]], dtype=np.float64)
    out["bound"] = _np(bound); out["truncation"] = np.array(t.truncation); out["edge"] = np.array(case["edge"])
    out["n_stratified"] = np.array(cfg["rendering"]["n_stratified"]); out["n_importance"] = np.array(cfg["rendering"]["n_importance"])
    out["log2_hash"] = np.array([cfg["grid"]["hash_size_sdf"], cfg["grid"]["hash_size_color"]])
    out["per_level_scale"] = np.array([grids[0].spec.per_level_scale, grids[1].spec.per_level_scale])
    out["variant"] = np.array("B" if cfg["grid"]["tcnn_network"] else "A")
    out["color_img"] = _np(col); out["depth_img"] = _np(dep); out["cam_pose"] = _np(cam_pose0); out["mask_mode"] = np.array(t.t_mask_mode)
    out["indices"] = _np([t_ for k, t_ in rec.draws if k == "randint"][-1])
    out["t_rand"] = _np([t_ for k, t_ in rec.draws if k == "rand"][-1])
    s = rec.samples[-1]
    for nm, o in zip(("rays_o", "rays_d", "depth", "color"), s["out"]):
        out["sample_out_" + nm] = _np(o)
    r = rec.renders[-1]
    out["render_rays_o"] = _np(r["rays_o"]); out["render_rays_d"] = _np(r["rays_d"]); out["render_gt_depth"] = _np(r["gt_depth"])
    for nm, o in zip(("term", "pixel_unc", "depth", "rgb", "sdf", "z_vals", "depth_unc"), r["out"]):
        out["ret_" + nm] = _np(o)
    out["loss"] = np.array(loss); out["pixel_unc"] = _np(punc)
    out["grad_T"] = _np(rec.steps[-1][0][0]); out["grad_R"] = _np(rec.steps[-1][1][0])
    os.makedirs(OUT, exist_ok=True)
    _save(name, out)
    print(name, "rays", r["gt_depth"].numel(), "loss", loss, "gradT", out["grad_T"], "gradR", out["grad_R"])


def gen_mesh_query(name, case):
    from src.utils.Mesher import Mesher
    cfg = _load_cfg(case)
    bound, grids, dec = _build_world(cfg, 120)
    m = object.__new__(Mesher)                                   # __init__ opens the dataset; the two methods need only these
    m.cfg = cfg; m.bound = bound; m.points_batch_size = 700      # several ragged chunks
    m.marching_cubes_bound = torch.from_numpy(np.array(cfg["mapping"]["marching_cubes_bound"]) * cfg["scale"])   # Mesher.py:56-57
    # numpy >= 2 hands torch 0-d tensors back from np.linspace as Tensors (the reference predates that): give linspace the
    # same float64 endpoints as python floats for the duration of the call -- the arithmetic is unchanged
    orig_linspace = np.linspace
    np.linspace = lambda a, b, n, **k: orig_linspace(float(a), float(b), int(n), **k)
    try:
        grid = m.get_grid_uniform(case["resolution"])
    finally:
        np.linspace = orig_linspace
    pts = grid["grid_points"]
    with torch.no_grad():
        ret = m.eval_points(pts, ([grids[0]], [grids[1]]), dec, "cpu")
    out = {}
    out["bound"] = _np(bound); out["mc_bound"] = _np(m.marching_cubes_bound); out["resolution"] = np.array(case["resolution"])
    out["log2_hash"] = np.array([cfg["grid"]["hash_size_sdf"], cfg["grid"]["hash_size_color"]])
    out["per_level_scale"] = np.array([grids[0].spec.per_level_scale, grids[1].spec.per_level_scale])
    out["variant"] = np.array("B" if cfg["grid"]["tcnn_network"] else "A")
    for a, nm in zip(grid["xyz"], "xyz"):
        out["axis_" + nm] = np.asarray(a)                        # float64 np.linspace, as the reference keeps them
    out["points"] = _np(pts); out["sdf"] = _np(ret[:, 3]); out["rgb"] = _np(ret[:, :3])
    os.makedirs(OUT, exist_ok=True)
    _save(name, out)
    print(name, "points", pts.shape[0], "dims", [len(a) for a in grid["xyz"]], "outside", int((ret[:, 3] == -1).sum()))


def gen_covisibility(name, case):
    """Runs the unmodified Mapper.keyframe_selection_LC and records its inputs, the torch.randint draw, percent_inside (the
    argument of its torch.argmax) and the returned selection."""
    from src.Mapper import Mapper
    cfg = _load_cfg(case)
    H, W = case["H"], case["W"]
    cam = cfg["cam"]
    n_kf = case["n_kf"]
    frames, dirs = _frames(cfg, case, n_kf + 1, seed=17)
    m = object.__new__(Mapper)
    m.cfg = cfg; m.device = "cpu"
    m.H, m.W, m.fx, m.fy, m.cx, m.cy = H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    m.LC = cfg["mapping"]["LC"]; m.LC_cnt = torch.zeros(1).int(); m.tracking_back = torch.tensor([0])
    m.activated_mapping_mode = False
    torch.manual_seed(23)
    m.keyframe_list = [4 * k for k in range(n_kf)]
    est = torch.zeros(4 * (n_kf + 1) + 1, 4, 4)
    for k in range(n_kf):
        c = frames[k][2].clone(); c[:3, 3] += 0.01 * torch.randn(3)
        est[4 * k] = c
    m.estimate_c2w_list = est
    col, dep, c2w = frames[n_kf]
    rec = Recorder(); rec.install()
    percent = []
    orig_argmax = torch.argmax

    def argmax(t, *a, **k):
        percent.append(t.detach().clone()); return orig_argmax(t, *a, **k)
    torch.argmax = argmax
    try:
        torch.manual_seed(29)
        sel = m.keyframe_selection_LC(n_kf - 2, 4 * n_kf, col, dep, c2w, 4)
    finally:
        torch.argmax = orig_argmax
        rec.uninstall()
    out = {"meta_H_W_fx_fy_cx_cy": np.array([H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"]], dtype=np.float64),
           "color_img": _np(col), "depth_img": _np(dep), "c2w": _np(c2w), "keyframe_c2ws": _np(torch.stack([est[i] for i in m.keyframe_list])),
           "indices": _np([t for k, t in rec.draws if k == "randint"][-1]), "percent_inside": _np(percent[0]), "selected": np.array(sel),
           "num_samples": np.array(8), "num_rays": np.array(50), "edge": np.array(20)}
    s_ = rec.samples[-1]
    for nm, o in zip(("rays_o", "rays_d", "depth", "color"), s_["out"]):
        out["sample_out_" + nm] = _np(o)
    _save(name, out)
    print(name, "keyframes", n_kf, "percent_inside", _np(percent[0]).round(3), "selected", sel)


def gen_render_img(name, case):
    from src.utils.Renderer import Renderer
    cfg = _load_cfg(case)
    bound, grids, dec = _build_world(cfg, 80)
    H, W = case["H"], case["W"]
    cam = cfg["cam"]
    frames, _ = _frames(cfg, case, 3, seed=13)
    col, dep, c2w = frames[2]
    fake = types.SimpleNamespace(bound=bound, device="cpu", H=H, W=W, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"])
    renderer = Renderer(cfg, fake, ray_batch_size=case["ray_batch"])
    rec = Recorder(); rec.install()
    try:
        torch.manual_seed(5)
        ret = renderer.render_img(([grids[0]], [grids[1]]), dec, c2w, cfg["model"]["truncation"], "cpu", gt_depth=dep)
    finally:
        rec.uninstall()
    out = {}
    out["meta_H_W_fx_fy_cx_cy"] = np.array([H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"]], dtype=np.float64)
    out["bound"] = _np(bound); out["truncation"] = np.array(cfg["model"]["truncation"]); out["ray_batch"] = np.array(case["ray_batch"])
    out["n_stratified"] = np.array(cfg["rendering"]["n_stratified"]); out["n_importance"] = np.array(cfg["rendering"]["n_importance"])
    out["log2_hash"] = np.array([cfg["grid"]["hash_size_sdf"], cfg["grid"]["hash_size_color"]])
    out["per_level_scale"] = np.array([grids[0].spec.per_level_scale, grids[1].spec.per_level_scale])
    out["variant"] = np.array("B" if cfg["grid"]["tcnn_network"] else "A")
    out["c2w"] = _np(c2w); out["depth_img"] = _np(dep)
    # torch.rand draws in consumption order: per chunk (n_valid,S) then, if the chunk has holes, (n0,n_strat), (n0,n_imp)
    rands = [t for k, t in rec.draws if k == "rand"]
    out["n_draws"] = np.array(len(rands))
    out["draw_shapes"] = np.array([list(t.shape) for t in rands], dtype=np.int64)
    out["draws_flat"] = np.concatenate([_np(t).reshape(-1) for t in rands])
    for nm, o in zip(("depth", "color", "term", "pixel_unc", "depth_unc"), ret):
        out["ret_" + nm] = _np(o)
        out["dtype_" + nm] = np.array(str(o.dtype))
    os.makedirs(OUT, exist_ok=True)
    _save(name, out)
    print(name, "pixels", H * W, "holes", int((dep == 0).sum()), "draws", len(rands), "mean depth", float(ret[0].mean()))


class _Mesh:
    """Stand-in for the trimesh.Trimesh container methods cull_mesh.py calls (trimesh is not in this image): vertices,
    faces, update_faces(mask), remove_unreferenced_vertices(), process(), export()."""

    def __init__(self, vertices, faces):
        self.vertices = np.asarray(vertices, dtype=np.float64); self.faces = np.asarray(faces, dtype=np.int64)
        self.face_masks = []

    def update_faces(self, mask):
        self.face_masks.append(np.asarray(mask).copy())
        self.faces = self.faces[np.asarray(mask)]

    def remove_unreferenced_vertices(self):
        ref = np.zeros(len(self.vertices), dtype=bool)
        ref[self.faces.reshape(-1)] = True
        remap = np.cumsum(ref) - 1
        self.vertices = self.vertices[ref]; self.faces = remap[self.faces]

    def process(self, validate=False):
        pass                      # trimesh's vertex merge: the test mesh has no duplicate vertices

    def export(self, path):
        self.exported = path


class _Hull:
    """mesh_bound of Mesher.get_bound_from_frames: a closed convex hull; contains() as the half-space test."""

    def __init__(self, planes):
        self.planes = planes

    def contains(self, pts):
        from oracle import cull_ref
        return cull_ref.inside_hull(pts, self.planes)[0]


def gen_cull(name, case):
    """Runs the unmodified cull_mesh (eval_rec True and False) and cull_out_bound_mesh.  The per-vertex visibility is read
    off one degenerate face (i,i,i) per vertex appended to the face list: update_faces() receives ~whole_mask[i] for it."""
    import scipy.spatial
    import src.tools.cull_mesh as CM
    from oracle import cull_ref, mc_ref
    from oracle import scene as syn
    cfg = _load_cfg(case)
    H, W = case["H"], case["W"]
    cam = cfg["cam"]
    K = case["n_frames"]
    frames, dirs = _frames(cfg, case, 12 * K, seed=31)
    frames = frames[::12]                                              # poses 48 trajectory steps apart
    torch.manual_seed(37)
    est = []
    for _, _, c2w in frames:
        c = c2w.clone(); c[:3, 3] += 0.01 * torch.randn(3); est.append(c)
    # the mesh: marching cubes (oracle) of the analytic room on a coarse grid, PLY precision (fp32)
    room = syn.AnalyticRoom(cfg["mapping"]["bound"])
    b = np.asarray(cfg["mapping"]["bound"], dtype=np.float64)
    axes = [np.arange(lo - 0.2, hi + 0.2, case["voxel"], dtype=np.float32) for lo, hi in b]
    gx, gy, gz = np.meshgrid(axes[0], axes[1], axes[2], indexing="xy")            # (ny, nx, nz) like the Mesher's volume
    vol = room.sdf(torch.from_numpy(np.stack([gx, gy, gz], -1))).numpy()
    verts, faces, _ = mc_ref.marching_cubes(vol, 0.0, [a[0] for a in axes], [case["voxel"]] * 3)
    V = len(verts)
    faces_aug = np.concatenate([faces, np.repeat(np.arange(V)[:, None], 3, axis=1)], axis=0)
    reader = [(k, frames[k][0], frames[k][1], frames[k][2], dirs) for k in range(K)]
    orig = (CM.get_dataset, getattr(CM.trimesh, "load", None))
    CM.get_dataset = lambda cfg_, args_, scale_, device="cpu": reader
    out = {"meta_H_W_fx_fy_cx_cy": np.array([H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"]], dtype=np.float64),
           "truncation": np.array(cfg["model"]["truncation"]), "verts": verts, "faces": faces,
           "c2ws": _np(torch.stack(est)), "depths": _np(torch.stack([f[1] for f in frames]))}
    try:
        for tag, eval_rec in (("rec", True), ("vis", False)):
            mesh = _Mesh(verts, faces_aug)
            CM.trimesh.load = lambda f, process=False, _m=mesh: _m
            CM.cull_mesh("mesh.ply", cfg, None, "cpu", eval_rec, estimate_c2w_list=est)
            keep = mesh.face_masks[0]
            out[f"seen_{tag}"] = keep[len(faces):]                      # per vertex: ~whole_mask
            out[f"face_keep_{tag}"] = keep[:len(faces)]
            assert mesh.exported == "mesh_culled.ply"
            m2 = _Mesh(verts, faces)                                     # the culled mesh without the probe faces
            m2.update_faces(keep[:len(faces)]); m2.remove_unreferenced_vertices()
            out[f"culled_verts_{tag}"] = m2.vertices.astype(np.float32); out[f"culled_faces_{tag}"] = m2.faces
        # convex bound: hull of the camera centres and a subset of back-projected depth points, scaled about its centre
        # (what get_bound_from_frames assembles from the TSDF fusion; Mesher.py:117-131): with a handful of frames it cuts the mesh
        pts = [c[:3, 3].numpy() for c in est]
        for (_, dep, _), c in zip(frames, est):
            d = (dirs.reshape(-1, 3) @ c[:3, :3].T) * dep.reshape(-1, 1) + c[:3, 3]
            pts.append(d[dep.reshape(-1) > 0][::17].numpy())
        pts = np.concatenate([np.atleast_2d(p) for p in pts], axis=0).astype(np.float64)
        ctr = pts.mean(axis=0)
        pts = ctr + cfg["meshing"]["mesh_bound_scale"] * (pts - ctr)
        hull = scipy.spatial.ConvexHull(pts)
        hv = pts[hull.vertices]
        remap = -np.ones(len(pts), dtype=np.int64); remap[hull.vertices] = np.arange(len(hull.vertices))
        hf = remap[hull.simplices]
        planes = cull_ref.hull_planes(hv, hf)
        mesh = _Mesh(verts, faces_aug)
        ret = CM.cull_out_bound_mesh(mesh, _Hull(planes), cfg, None, "cpu", est)
        keep = mesh.face_masks[0]
        out["hull_verts"] = hv; out["hull_faces"] = hf
        out["inside_hull"] = keep[len(faces):]; out["face_keep_hull"] = keep[:len(faces)]
        m2 = _Mesh(verts, faces)
        m2.update_faces(keep[:len(faces)]); m2.remove_unreferenced_vertices()
        out["culled_verts_hull"] = m2.vertices.astype(np.float32); out["culled_faces_hull"] = m2.faces
    finally:
        CM.get_dataset = orig[0]
        if orig[1] is not None:
            CM.trimesh.load = orig[1]
    _save(name, out)
    print(name, "vertices", V, "faces", len(faces), "seen rec/vis", int(out["seen_rec"].sum()), int(out["seen_vis"].sum()),
          "inside hull", int(out["inside_hull"].sum()), "hull faces", len(hf))


def gen_eval_rendering(name, case):
    """Runs the unmodified eval_rendering over a short sequence of tiny frames with the reference's own Renderer (LPIPS and
    MS-SSIM, third-party networks, are stand-ins that return 0).  Records, per evaluated frame, what the metrics are computed
    from (dataset colour -- float64 as datasets.py:87-91 produces it -- and depth, rendered colour and depth), the argument
    of its torch.log10 (the frame's mse, full precision), the outputs of its torch.abs (the depth residuals) and the result
    line it appends to output.txt."""
    import json
    import tempfile
    import src.tools.eval_recon as ER
    from src.utils.Renderer import Renderer
    cfg = _load_cfg(case)
    bound, grids, dec = _build_world(cfg, 90)
    H, W = case["H"], case["W"]
    cam = cfg["cam"]
    n_img = case["n_img"]
    frames, dirs = _frames(cfg, case, n_img, seed=41)
    fake = types.SimpleNamespace(bound=bound, device="cpu", H=H, W=W, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"])
    renderer = Renderer(cfg, fake, ray_batch_size=case["ray_batch"])
    reader = [(k, frames[k][0].double(), frames[k][1], frames[k][2], dirs) for k in range(n_img)]
    torch.manual_seed(43)
    est = []
    for _, _, c2w in frames:
        c = c2w.clone(); c[:3, 3] += 0.005 * torch.randn(3); est.append(c)
    rendered, mses, residuals = [], [], []
    orig_render, orig_log10, orig_abs = renderer.render_img, torch.log10, torch.abs

    def render(*a, **k):
        out = orig_render(*a, **k); rendered.append([o.detach().clone() for o in out]); return out

    def log10(t):
        mses.append(t.detach().clone()); return orig_log10(t)

    def tabs(t):
        out = orig_abs(t); residuals.append(out.detach().clone()); return out
    renderer.render_img = render; torch.log10 = log10; torch.abs = tabs
    outdir = tempfile.mkdtemp(prefix="usl_evalr_")
    try:
        torch.manual_seed(3)
        ER.eval_rendering(cfg, n_img, reader, est, renderer, ([grids[0]], [grids[1]]), dec, cfg["model"]["truncation"], outdir, "cpu")
    finally:
        torch.log10 = orig_log10; torch.abs = orig_abs
    result = json.loads(open(os.path.join(outdir, "output.txt")).read().strip().splitlines()[0])
    idxs = list(range(0, n_img, 5))
    assert len(rendered) == len(idxs) == len(mses)
    res = [r for r in residuals if r.dim() == 1][-len(idxs):]              # the depth residuals (1-D, after the boolean index)
    out = {"frames": np.array(idxs), "H_W": np.array([H, W]),
           "gt_color": np.stack([_np(reader[i][1]) for i in idxs]), "gt_depth": np.stack([_np(reader[i][2]) for i in idxs]),
           "depth": np.stack([_np(r[0]) for r in rendered]), "color": np.stack([_np(r[1]) for r in rendered]),
           "mse": np.array([float(m) for m in mses], dtype=np.float64), "mse_dtype": np.array(str(mses[0].dtype)),
           "depth_l1": np.array([float(r.double().mean()) for r in res], dtype=np.float64),
           "avg_psnr": np.array(result["avg_psnr"]), "depth_l1_render": np.array(result["depth_l1_render"])}
    _save(name, out)
    print(name, "frames", idxs, "mse", out["mse"], "psnr", result["avg_psnr"], "depth_l1", result["depth_l1_render"], out["mse_dtype"])


def main():
    _setup_paths()
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        # 8 threads everywhere but where a case says otherwise: at map_scannet_k23's size torch's multi-threaded scatter-add (the
        # autograd of the table gathers) sums in a run-dependent order, so that case runs on one thread to stay reproducible
        torch.set_num_threads(case.get("threads", 8))
        if name.startswith("map"):
            gen_mapping(name, case)
        elif name.startswith("kf_covis"):
            gen_covisibility(name, case)
        elif name.startswith("img"):
            gen_render_img(name, case)
        elif name.startswith("mesh"):
            gen_mesh_query(name, case)
        elif name.startswith("cull"):
            gen_cull(name, case)
        elif name.startswith("evalr"):
            gen_eval_rendering(name, case)
        else:
            gen_tracking(name, case)


if __name__ == "__main__":
    main()
