"""oracle/scene.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Frozen copy of the analytic indoor scene (room + furniture SDF, sphere tracing, procedural albedo, Lissajous
trajectory) that oracle/gen_golden.py renders its tiny input frames from.  It lives here, not in the product
package, so that edits to uni-slam_b200/synthetic.py can never change what the committed generator produces:
`python -m oracle.gen_golden` must reproduce tests/golden/*.npz bit for bit (tests/test_golden_reproduce.py).
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


