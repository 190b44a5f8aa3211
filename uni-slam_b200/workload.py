"""Synthetic BASELINE workloads (BASELINE.json configs 2-5): the tensors one Mapper / Tracker iteration
consumes, built with the reference's own constants (configs/Replica/*.yaml, configs/ScanNet/*.yaml,
src/Mapper.py:315-356,379-393,528-541).  Pure torch plumbing, device-agnostic, so the same builder feeds
the CUDA path and (moved to CPU) the oracle in bench.py's cpu_baseline leg."""
import math
from dataclasses import dataclass
from typing import List

import numpy as np
import torch

from . import synthetic as syn


@dataclass
class MappingWorkload:
    cfg: syn.SceneCfg
    bound: torch.Tensor            # (3,2) after load_bound
    per_level_scale: float
    c2ws: torch.Tensor             # (K,4,4) estimated poses of the window (first = fixed)
    cam_poses: torch.Tensor        # (K-1,7) joint-opt parameters (matrix_to_cam_pose of c2ws[1:])
    depths: torch.Tensor           # (K,P)
    colors: torch.Tensor           # (K,P,3)
    dirs_cam: torch.Tensor         # (K,P,3)
    n_main: int                    # pixels per frame of the main get_samples_all call (mapping.pixels // K)
    n_recent: int                  # 200 px x last 10 frames when > 20 keyframes (Mapper.py:385-393)
    cur_frame: tuple               # (color (H,W,3), depth (H,W), gt c2w) of the current frame, for tracking

    @property
    def K(self):
        return self.c2ws.shape[0]

    @property
    def P(self):
        return self.depths.shape[1]

    @property
    def n_rays(self):
        return self.K * self.n_main + (10 * self.n_recent if self.n_recent else 0)

    @property
    def S(self):
        return self.cfg.n_stratified + self.cfg.n_importance

    def draw(self, gen=None):
        """The torch.randint / torch.rand draws of one iteration (host-code side of the RNG contract)."""
        dev = self.depths.device
        R, S = self.n_rays, self.S
        idx_main = torch.randint(self.P, (self.K * self.n_main,), device=dev, generator=gen)
        idx_recent = torch.randint(self.P, (10 * self.n_recent,), device=dev, generator=gen) if self.n_recent else None
        t_rand = torch.rand((R, S), device=dev, generator=gen)
        t_uni = torch.rand((R, self.cfg.n_stratified), device=dev, generator=gen)
        u_pdf = torch.rand((R, self.cfg.n_importance), device=dev, generator=gen)
        return idx_main, idx_recent, t_rand, t_uni, u_pdf

    def alloc_draws(self):
        """Static input buffers for the iteration's draws, laid out so that ONE torch.randint-style fill and ONE
        torch.rand-style fill produce all of them (the fused driver indexes draws by ray slot, so the five tensors
        the reference draws separately can live in two flat buffers).  Returns (flat_idx, flat_u, views) with
        views = (idx_main, idx_recent|None, t_rand, t_uni, u_pdf) in draw() order."""
        dev = self.depths.device
        R, S, ns, ni = self.n_rays, self.S, self.cfg.n_stratified, self.cfg.n_importance
        n_main, n_rec = self.K * self.n_main, (10 * self.n_recent if self.n_recent else 0)
        flat_idx = torch.zeros((n_main + n_rec,), device=dev, dtype=torch.int64)
        flat_u = torch.zeros((R * (S + ns + ni),), device=dev, dtype=torch.float32)
        views = (flat_idx[:n_main], flat_idx[n_main:] if n_rec else None, flat_u[:R * S].view(R, S),
                 flat_u[R * S:R * (S + ns)].view(R, ns), flat_u[R * (S + ns):].view(R, ni))
        return flat_idx, flat_u, views

    def batches(self, idx_main, idx_recent):
        b = [(self.c2ws, self.depths, self.colors, self.dirs_cam, idx_main, self.n_main, 0)]
        if self.n_recent:
            K = self.K
            b.append((self.c2ws[K - 10:], self.depths[K - 10:], self.colors[K - 10:], self.dirs_cam[K - 10:], idx_recent,
                      self.n_recent, K - 10))
        return b


def _matrix_to_cam_pose(c2w: torch.Tensor) -> torch.Tensor:
    """Host-side pose helper (common.py:182-194) -- only used to seed the synthetic workload."""
    from .compat_pytorch3d.pytorch3d.transforms import matrix_to_quaternion
    return torch.cat([matrix_to_quaternion(c2w[:, :3, :3]), c2w[:, :3, 3]], dim=-1)


def build_mapping_workload(cfg: syn.SceneCfg, device, n_keyframes: int = 21, seed: int = 1, frame_stride: int = None,
                           scale_hw: float = 1.0) -> MappingWorkload:
    """Window of n_keyframes keyframes + the current frame, each stored as a 10 % pixel subset."""
    frame_stride = frame_stride or cfg.map_every
    seq = syn.SyntheticSequence(cfg, n_frames=200, device=device, seed=seed, scale_hw=scale_hw)
    g = torch.Generator(device="cpu").manual_seed(seed)
    K = n_keyframes + 1
    H, W = seq.cam.H, seq.cam.W
    P = int(H * W * 0.1)
    dirs_full = seq.dirs.reshape(-1, 3)
    c2ws, depths, colors, dirs = [], [], [], []
    cur = None
    for k in range(K):
        col, dep, c2w = seq.frame(k * frame_stride)
        ind = torch.randperm(H * W, generator=g)[:P].to(device)
        est = c2w.clone()
        if k > 0:   # estimated poses carry a little tracking noise
            est[:3, 3] += 0.005 * torch.randn(3, generator=g).to(device)
        c2ws.append(est); depths.append(dep.reshape(-1)[ind]); colors.append(col.reshape(-1, 3)[ind]); dirs.append(dirs_full[ind])
        cur = (col, dep, c2w)
    c2ws = torch.stack(c2ws).contiguous()
    bound = syn.load_bound(cfg.bound_yaml)
    res = syn.grid_resolution(bound, cfg.voxel)
    pls = float(np.exp2(np.log2(res / 16) / 15))
    n_main = cfg.map_pixels // K
    n_recent = 200 if n_keyframes > 20 else 0                      # len(keyframe_list) > 20 (Mapper.py:385)
    return MappingWorkload(cfg, bound, pls, c2ws, _matrix_to_cam_pose(c2ws[1:]).contiguous(), torch.stack(depths).contiguous(),
                           torch.stack(colors).contiguous(), torch.stack(dirs).contiguous(), n_main, n_recent, cur)


def init_field_tensors(cfg: syn.SceneCfg, wl_bound: torch.Tensor, pls: float, device, seed: int = 0):
    """Random-init hash tables (tcnn U(-1e-4,1e-4)) and decoders (framework default init), beta = 10."""
    from . import _lib as L
    from . import ops
    g = torch.Generator().manual_seed(seed)
    grids = [L.build_grid(16, cfg.log2_hash_sdf, 16, pls), L.build_grid(16, cfg.log2_hash_color, 16, pls)]
    tabs = [((torch.rand(gr.total_entries * 2, generator=g) * 2 - 1) * 1e-4).to(device) for gr in grids]
    if cfg.decoder_variant == "B":
        xav = lambda o, i: (torch.rand(o, i, generator=g) * 2 - 1) * math.sqrt(6.0 / (i + o))
        dec = [torch.cat([xav(16, 32).reshape(-1), xav(16, 16).reshape(-1)]).to(device) for _ in range(2)]
    else:
        def lin(o, i):
            k = 1.0 / math.sqrt(i)
            return [((torch.rand(o, i, generator=g) * 2 - 1) * k).to(device), ((torch.rand(o, generator=g) * 2 - 1) * k).to(device)]
        dec = lin(16, 32) + lin(16, 16) + lin(1, 16) + lin(16, 32) + lin(16, 16) + lin(3, 16)
    beta = torch.full((1,), 10.0, device=device)
    meta = ops.FieldMeta(grids[0], grids[1], cfg.decoder_variant, L.make_bound(wl_bound))
    return meta, tabs, dec, beta
