"""Synthetic Replica-/ScanNet-shaped RGB-D frames from an analytic indoor SDF scene.

Datasets are not available offline, so the frames the reference would
read through src/utils/datasets.py:76-138 (colour HxWx3 in [0,1], depth HxW in metres with
0 = hole, OpenGL-convention c2w 4x4, per-pixel camera dirs from get_camera_rays,
src/common.py:35-46) are generated here: an axis-aligned room a few cm inside the scene bound
with box / sphere / cylinder furniture, sphere-traced optical-axis depth (dir_z = -1, so
depth == ray parameter t), procedural albedo, a smooth Lissajous trajectory and random
zero-depth holes.  Pure torch, runs on CPU or CUDA (this is data plumbing, not the hot path).
"""
import math
from dataclasses import dataclass

import torch


@dataclass
class CameraCfg:
    H: int
    W: int
    fx: float
    fy: float
    cx: float
    cy: float


@dataclass
class SceneCfg:
    name: str
    cam: CameraCfg
    bound_yaml: tuple            # mapping.bound of the yaml
    voxel: float                 # grid.voxel_sdf == voxel_color
    log2_hash_sdf: int
    log2_hash_color: int
    n_stratified: int
    n_importance: int
    decoder_variant: str         # 'A' nn.Linear fp32 (tcnn_network False) / 'B' FullyFusedMLP restated
    track_iters: int
    map_iters: int
    map_every: int
    hash_lr: float
    lr_T: float
    lr_R: float
    hole_frac: float
    depth_quant: float           # 1/png_depth_scale quantisation step (0 = none)
    truncation: float = 0.06
    ignore_edge: int = 75
    track_pixels: int = 2000
    map_pixels: int = 4000


# configs/Replica/replica.yaml + room0.yaml ; configs/ScanNet/scannet.yaml + scene0000.yaml (after crop_edge)
REPLICA_ROOM0 = SceneCfg("replica_room0", CameraCfg(680, 1200, 600.0, 600.0, 599.5, 339.5),
                         ((-1.0, 7.0), (-1.3, 3.7), (-1.7, 1.4)), 0.01, 16, 19, 32, 8, "B",
                         8, 15, 4, 0.05, 0.002, 0.001, 0.02, 0.0)
SCANNET_0000 = SceneCfg("scannet_scene0000", CameraCfg(460, 620, 577.590698, 578.729797, 308.905426, 232.683609),
                        ((-0.1, 8.6), (-0.1, 8.9), (-0.3, 3.3)), 0.02, 16, 16, 48, 8, "A",
                        20, 20, 5, 0.02, 0.0005, 0.0025, 0.07, 0.001)
CONFIGS = {c.name: c for c in (REPLICA_ROOM0, SCANNET_0000)}


def load_bound(bound_yaml, bound_dividable=0.24, scale=1.0) -> torch.Tensor:
    """src/UNISLAM.py:205-222: hi rounded up so (hi-lo) is a multiple of 0.24 (fp32 arithmetic)."""
    bound = (torch.tensor(bound_yaml, dtype=torch.float64) * scale).float()
    bound[:, 1] = (((bound[:, 1] - bound[:, 0]) / bound_dividable).int() + 1) * bound_dividable + bound[:, 0]
    return bound


def grid_resolution(bound: torch.Tensor, voxel: float) -> int:
    """src/UNISLAM.py:192-199."""
    return int((bound[:, 1] - bound[:, 0]).max() / voxel)


def camera_dirs(cam: CameraCfg, device="cpu") -> torch.Tensor:
    """get_camera_rays, src/common.py:35-46 (OpenGL). (H,W,3) fp32."""
    i, j = torch.meshgrid(torch.arange(cam.W, dtype=torch.float32, device=device),
                          torch.arange(cam.H, dtype=torch.float32, device=device), indexing="xy")
    return torch.stack([(i - cam.cx) / cam.fx, -(j - cam.cy) / cam.fy, -torch.ones_like(i)], -1)


class AnalyticRoom:
    """SDF > 0 in free space, < 0 inside walls / furniture."""

    def __init__(self, bound_yaml, device="cpu", margin=0.12):
        b = torch.tensor(bound_yaml, dtype=torch.float32, device=device)
        lo, hi = b[:, 0] + margin, b[:, 1] - margin
        self.device = device
        self.c_room = 0.5 * (lo + hi)
        self.h_room = 0.5 * (hi - lo)
        ext = hi - lo
        f = lambda u, v, w: lo + ext * torch.tensor([u, v, w], device=device)
        # furniture: two boxes on the floor (z is up in our synthetic world's third axis), a sphere, a cylinder
        self.boxes = [(f(0.25, 0.30, 0.12), ext * torch.tensor([0.10, 0.14, 0.12], device=device)),
                      (f(0.72, 0.70, 0.20), ext * torch.tensor([0.08, 0.10, 0.20], device=device))]
        self.sphere = (f(0.55, 0.35, 0.30), float(ext.min()) * 0.16)
        self.cyl = (f(0.30, 0.75, 0.0), float(ext.min()) * 0.10, float(ext[2]) * 0.45)   # base centre, radius, height

    @staticmethod
    def _sd_box(p, c, h):
        q = (p - c).abs() - h
        return q.clamp(min=0).norm(dim=-1) + q.max(dim=-1)[0].clamp(max=0)

    def sdf(self, p: torch.Tensor) -> torch.Tensor:
        d = -self._sd_box(p, self.c_room, self.h_room)
        for c, h in self.boxes:
            d = torch.minimum(d, self._sd_box(p, c, h))
        sc, sr = self.sphere
        d = torch.minimum(d, (p - sc).norm(dim=-1) - sr)
        cc, cr, ch = self.cyl
        q = p - cc
        dr = q[..., :2].norm(dim=-1) - cr
        dz = (q[..., 2] - 0.5 * ch).abs() - 0.5 * ch
        dcyl = torch.stack([dr, dz], -1).clamp(min=0).norm(dim=-1) + torch.maximum(dr, dz).clamp(max=0)
        return torch.minimum(d, dcyl)

    def albedo(self, p: torch.Tensor) -> torch.Tensor:
        """Smooth + checker procedural colour in [0,1]."""
        s = 0.5 + 0.5 * torch.sin(p * torch.tensor([1.7, 2.3, 2.9], device=p.device) + torch.tensor([0.3, 1.1, 2.0], device=p.device))
        chk = ((torch.floor(p[..., 0] * 2.0) + torch.floor(p[..., 1] * 2.0) + torch.floor(p[..., 2] * 2.0)) % 2.0)
        return (0.65 * s + 0.30 * chk[..., None] * torch.tensor([0.9, 0.8, 0.6], device=p.device) + 0.03).clamp(0, 1)

    def trace(self, o: torch.Tensor, d: torch.Tensor, n_steps: int = 96, t_max: float = 20.0):
        """Sphere-trace rays p = o + d*t (d NOT normalised; dir_z=-1 => t is optical-axis depth).
        Returns t (0 where no hit) and hit points."""
        dn = d.norm(dim=-1)
        t = torch.zeros(o.shape[:-1], device=o.device)
        for _ in range(n_steps):
            p = o + d * t[..., None]
            t = t + self.sdf(p) / dn * 0.98
        p = o + d * t[..., None]
        hit = (self.sdf(p).abs() < 2e-3) & (t < t_max) & (t > 0)
        return torch.where(hit, t, torch.zeros_like(t)), p


def look_at_c2w(eye: torch.Tensor, target: torch.Tensor, up=(0.0, 0.0, 1.0)) -> torch.Tensor:
    """OpenGL camera (looks along -z, +y up) camera-to-world."""
    up = torch.tensor(up, dtype=torch.float32, device=eye.device)
    zc = eye - target
    zc = zc / zc.norm()
    xc = torch.linalg.cross(up, zc)
    xc = xc / xc.norm()
    yc = torch.linalg.cross(zc, xc)
    c2w = torch.eye(4, device=eye.device)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = xc, yc, zc, eye
    return c2w


def trajectory(room: AnalyticRoom, n_frames: int, period: int = 800) -> torch.Tensor:
    """Smooth Lissajous path in the middle of the room, gaze sweeping the walls. (n,4,4).
    One lap takes `period` frames: ~0.8 cm and ~0.4 deg per frame, hand-held-camera speed like Replica / ScanNet."""
    out = []
    for k in range(n_frames):
        s = 2 * math.pi * k / period
        off = torch.tensor([0.22 * math.sin(s), 0.20 * math.sin(2 * s + 0.5), 0.10 * math.sin(3 * s)], device=room.device)
        eye = room.c_room + room.h_room * off
        ang = 0.9 * s + 0.4
        tgt = room.c_room + room.h_room * torch.tensor([0.9 * math.cos(ang), 0.9 * math.sin(ang), -0.35 + 0.2 * math.sin(2 * s)], device=room.device)
        out.append(look_at_c2w(eye, tgt))
    return torch.stack(out)


class SyntheticSequence:
    """frame(k) -> (color (H,W,3) fp32, depth (H,W) fp32, c2w (4,4)); same tuple layout the
    reference's datasets return (datasets.py:138) minus the index / rays_d (see camera_dirs)."""

    def __init__(self, cfg: SceneCfg, n_frames: int = 200, device="cpu", seed: int = 1, scale_hw: float = 1.0):
        self.cfg = cfg
        cam = cfg.cam
        if scale_hw != 1.0:   # reduced-resolution frames for CPU-sized tests (intrinsics scaled consistently)
            cam = CameraCfg(int(cam.H * scale_hw), int(cam.W * scale_hw), cam.fx * scale_hw, cam.fy * scale_hw,
                            cam.cx * scale_hw, cam.cy * scale_hw)
        self.cam = cam
        self.device = device
        self.room = AnalyticRoom(cfg.bound_yaml, device)
        self.poses = trajectory(self.room, n_frames)
        self.dirs = camera_dirs(cam, device)
        self.seed = seed
        self.n_frames = n_frames

    def render_pixels(self, c2w: torch.Tensor, dirs_cam: torch.Tensor):
        d = torch.sum(dirs_cam[..., None, :] * c2w[:3, :3], -1)
        o = c2w[:3, 3].expand(d.shape)
        t, p = self.room.trace(o, d)
        col = self.room.albedo(p)
        col = torch.where((t > 0)[..., None], col, torch.zeros_like(col))
        return col, t

    def frame(self, k: int):
        c2w = self.poses[k]
        col, dep = self.render_pixels(c2w, self.dirs)
        g = torch.Generator(device="cpu").manual_seed(self.seed * 100003 + k)
        H, W = dep.shape
        # holes: coarse random blocks (sensor drop-outs) covering ~hole_frac of the image
        bh, bw = max(H // 20, 1), max(W // 20, 1)
        blocks = (torch.rand(bh, bw, generator=g) < self.cfg.hole_frac).to(self.device)
        mask = blocks.repeat_interleave(-(-H // bh), 0)[:H].repeat_interleave(-(-W // bw), 1)[:, :W]
        dep = torch.where(mask, torch.zeros_like(dep), dep)
        if self.cfg.depth_quant > 0:
            dep = torch.round(dep / self.cfg.depth_quant) * self.cfg.depth_quant
        return col.contiguous(), dep.contiguous(), c2w
