"""Fused per-iteration drivers: one Mapper iteration / one Tracker iteration / the dense SDF query,
as fixed sequences of C-ABI launches on preallocated buffers (no host sync, no boolean-index
compaction, CUDA-graph capturable).

They compute what src/Mapper.py:366-445 (optimize_mapping loop body up to loss.backward()),
src/Tracker.py:149-244 (optimize_tracking up to loss.backward()) and src/utils/Mesher.py:134-227
(eval_points over get_grid_uniform) compute, and leave the gradients in ``.grad``-compatible buffers
so the host code's own torch.optim.Adam can step on them.

RNG contract: the torch.randint / torch.rand draws stay on the host side (reference behaviour);
because rays are masked instead of compacted, the draws are indexed by ray slot: t_rand is (R,S),
t_rand_uni (R,n_stratified), u_pdf (R,n_importance).  (The drop-in ``modules.Renderer`` keeps the
reference's compacted draw shapes instead.)
"""
import os
from ctypes import byref
from typing import List, Optional, Sequence

import torch

from . import _lib as L
from . import ops
from ._lib import call, cptr, ptr, stream


def _mask_mode(mask_mode: str, original: int) -> int:
    """cfg['m_mask_mode'] / cfg['t_mask_mode'] -> usl_loss_args_t.mode."""
    if mask_mode == "original":
        return original
    if mask_mode == "no_mask":
        return 2
    raise ValueError(f"mask_mode must be 'original' or 'no_mask', got {mask_mode!r}")


class _Profiled:
    """Optional per-kernel CUDA-event timing on the launching stream (bench.py's roofline leg)."""
    profile = False

    def _init_prof(self):
        self.events = {}          # kernel name -> list of (start, stop) events

    def _call(self, name, *args):
        if not self.profile:
            return call(name, *args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call(name, *args)
        e1.record()
        self.events.setdefault(name, []).append((e0, e1))

    def kernel_ms(self):
        """Mean duration per launch (ms) of every profiled entry point; call after a synchronize."""
        return {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in self.events.items()}


class _FieldState:
    """Tables + decoder tensors + beta, their packed descriptor and persistent gradient buffers."""

    def __init__(self, meta: ops.FieldMeta, sdf_table, rgb_table, dec: Sequence[torch.Tensor], beta: torch.Tensor,
                 with_grads: bool, max_frames: int = 1, grad_alloc=None):
        self.meta, self.sdf_table, self.rgb_table, self.dec, self.beta = meta, sdf_table, rgb_table, list(dec), beta
        self.field = meta.pack(sdf_table, rgb_table, self.dec)
        if with_grads:
            # every gradient lives in ONE flat buffer [colour table | sdf table | decoders | beta | poses]: a single memset
            # clears it and (multi-GPU) the exchange kernels sum it.  The colour table (87 % of the bytes) comes first so that
            # its exchange -- one contiguous range -- can start while the sdf half of the backward still runs, and the rest is
            # one more contiguous range.  Table sizes are multiples of 16 floats, so the 16-byte vector atomics stay aligned.
            n_scratch = ops.bwd_scratch_floats(self.field)
            sizes = [rgb_table.numel(), sdf_table.numel()] + [t.numel() for t in self.dec] + [1, max_frames * 7]   # ..., beta, d pose
            pad = (-sum(sizes)) % 32                                  # keep the scratch block 128-byte aligned (16-byte vector atomics)
            sizes += [pad, max(n_scratch, 4), L.LOSS_SLOTS, max_frames * 12]   # + loss accumulators + d c2w: one memset clears all
            # grad_alloc(numel) -> zero-filled fp32 tensor: multi-GPU runs hand out peer-mapped (symmetric) memory here so that
            # the exchange kernels of collective.cu can read / write every rank's buffer directly (parallel.PeerGroup)
            self.g_all = grad_alloc(sum(sizes)) if grad_alloc is not None else torch.zeros(sum(sizes), device=sdf_table.device, dtype=torch.float32)
            views, o = [], 0
            for s_ in sizes:
                views.append(self.g_all[o:o + s_])
                o += s_
            self.g_rgb_table, self.g_sdf_table = views[0], views[1]
            self.g_dec = [v.view(t.shape) for v, t in zip(views[2:-6], self.dec)]
            self.g_beta = views[-6]
            self.d_pose = views[-5].view(max_frames, 7)             # inside the gradient block: one all-reduce covers it
            self.scratch = views[-3]                                 # replicated coarse levels (usl_field_bwd workspace)
            self.acc = views[-2]
            self.d_c2w = views[-1].view(max_frames, 12)
            self.n_grad = sum(sizes[:-4])
            self.n_grad_padded = sum(sizes[:-3])                     # + the zero pad: a multiple of 32 floats (16-byte vector exchange)
            self.g_grads = self.g_all[:self.n_grad]                  # what a multi-GPU all-reduce must sum
            self.g_flat = self.g_all[sizes[0] + sizes[1]:self.n_grad]   # decoder + beta + pose gradients
            self.g_mlp = meta.pack_grads(self.g_dec)

    def repack(self):
        self.field = self.meta.pack(self.sdf_table, self.rgb_table, self.dec)

    def refresh(self, sdf_table=None, rgb_table=None, dec=None, beta=None):
        """Re-pack the descriptor at the top of every run(): a few ctypes stores, and an integrator that rebinds a table,
        a decoder tensor or beta (Tracker.update_params_from_mapping, Tracker.py:257-267; load_state_dict) is picked up
        instead of leaving the kernels on stale memory."""
        if sdf_table is not None:
            self.sdf_table = sdf_table
        if rgb_table is not None:
            self.rgb_table = rgb_table
        if dec is not None:
            self.dec = list(dec)
        if beta is not None:
            self.beta = beta
        for t, nm in ((self.sdf_table, "sdf table"), (self.rgb_table, "colour table"), (self.beta, "beta")):
            cptr(t, torch.float32, None, nm)
        self.field = self.meta.pack(self.sdf_table, self.rgb_table, self.dec)


class MappingStep(_Profiled):
    """One mapping iteration (sample -> prefilter -> z-sample -> render -> loss -> backward)."""

    def __init__(self, meta: ops.FieldMeta, sdf_table, rgb_table, dec, beta, *, n_stratified, n_importance, truncation,
                 weights=(5.0, 200.0, 10.0, 0.1, 5.0), max_rays: int, max_frames: int = 1, perturb: bool = True,
                 mask_mode: str = "original", grad_alloc=None):
        dev = sdf_table.device
        self.fs = _FieldState(meta, sdf_table, rgb_table, dec, beta, with_grads=True, max_frames=max_frames, grad_alloc=grad_alloc)
        self.zs = ops.ZSampler(n_stratified, n_importance, truncation, dev)
        self.S = self.zs.S
        self.perturb = perturb
        self.loss_args = ops.make_loss_args(truncation, weights[0], weights[1], weights[2], weights[3], weights[4],
                                            _mask_mode(mask_mode, 0))                       # cfg['m_mask_mode'], Mapper.py:94
        R, S = max_rays, self.S
        f32 = dict(device=dev, dtype=torch.float32)
        self.max_rays = R
        self.rays_o = torch.empty((R, 3), **f32); self.rays_d = torch.empty((R, 3), **f32)
        self.gt_depth = torch.empty((R,), **f32); self.gt_color = torch.empty((R, 3), **f32)
        self.dirs = torch.empty((R, 3), **f32); self.frame_id = torch.empty((R,), device=dev, dtype=torch.int32)
        self.valid = torch.empty((R,), device=dev, dtype=torch.uint8); self.mask = torch.empty((R,), device=dev, dtype=torch.uint8)
        self.z = torch.zeros((R, S), **f32)
        self.raw = torch.empty((R, S, 4), **f32); self.feat = torch.empty((ops.stash_floats(R * S),), **f32)   # stash: features + hidden pre-activations + clamped coordinates
        self.jac = torch.empty((12 * R * S,), **f32)        # component-major [12][n] of THIS call's n = R*S
        self.term = torch.empty((R,), **f32); self.punc = torch.empty((R,), **f32); self.depth = torch.empty((R,), **f32)
        self.rgb = torch.empty((R, 3), **f32); self.dunc = torch.empty((R,), **f32)
        self.acc = self.fs.acc; self.loss = torch.zeros((1,), **f32)
        self.d_raw = torch.empty((R, S, 4), **f32)
        self.d_rays_o = torch.empty((R, 3), **f32); self.d_rays_d = torch.empty((R, 3), **f32)
        self.d_c2w = self.fs.d_c2w; self.d_pose = self.fs.d_pose
        self.pdf_inds = None      # set record_pdf_inds(True) to keep sample_pdf's searchsorted indices (parity inspection)
        self.n_rays = 0
        self.acc_hook = None      # multi-GPU: called with self.acc between loss_fwd and loss_bwd (all-reduce of sums/counts)
        self.rgb_grads_hook = None  # multi-GPU: called with the colour-table gradient as soon as its half of field_bwd is queued
        self.bwd_leave_room = False # with the hook: the sdf half leaves one CTA slot per SM to the exchange kernel running beside it
        # independent pieces (gradient zero-fill; pose-gradient reduction) run on a side stream next to the critical path
        self.side_branches = True       # set False to run the zero-fill and the pose reduction on the main stream (ablation)
        self._side = None
        self._init_prof()

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream()
        return self._side

    def record_pdf_inds(self, on: bool = True):
        """Keep the torch.searchsorted indices of the no-depth branch (common.py:70) in self.pdf_inds (R, n_importance) int64."""
        self.pdf_inds = torch.full((self.max_rays, self.zs.n_importance), -1, device=self.z.device, dtype=torch.int64) if on else None

    # gradients, in the order Mapper.create_optimizer groups the parameters (Mapper.py:111-139)
    @property
    def grads(self):
        return dict(dec=self.fs.g_dec, beta=self.fs.g_beta, sdf_table=self.fs.g_sdf_table, rgb_table=self.fs.g_rgb_table)

    def run(self, batches, t_rand, t_rand_uni=None, u_pdf=None, cam_poses: Optional[torch.Tensor] = None,
            c2w_fixed: Optional[torch.Tensor] = None, has_holes: bool = True, ray_range=None):
        """batches: list of (c2ws|None, depths (K,P), colors (K,P,3), dirs_cam (K,P,3), indices (K*n,), n, frame_base).
        When cam_poses (K-1,7) is given (joint_opt, Mapper.py:372-376) the camera matrices are rebuilt on
        device as cat(c2w_fixed[None], pose_to_matrix(cam_poses)) and pose gradients land in self.d_pose.
        ray_range = (begin, end): process only that contiguous slice of the batch's global ray list (multi-GPU strong
        scaling, SURVEY 8e: every rank holds the same draws and takes its slice); the draws stay indexed by GLOBAL ray slot."""
        st = stream()
        fs, S = self.fs, self.S
        fs.refresh()
        joint = cam_poses is not None
        K = (cam_poses.shape[0] + 1) if joint else 0
        max_frames = self.d_pose.shape[0]
        if joint and K > max_frames:
            raise ValueError(f"MappingStep: window of {K} frames, gradient buffers sized for max_frames={max_frames}")
        if joint and (c2w_fixed is None or c2w_fixed.numel() < 12):
            raise ValueError("MappingStep: joint_opt needs c2w_fixed (4,4), the first keyframe's fixed pose (Mapper.py:360,374)")
        fork = self.side_branches and not self.profile
        cur = torch.cuda.current_stream()
        side = self._side_stream() if fork else None
        # gradients, replica scratch, loss accumulators, d c2w: one memset -- first needed by usl_loss_fwd, so it runs
        # beside ray set-up / z-sampling instead of in front of them
        if fork:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                fs.g_all.zero_()
        else:
            fs.g_all.zero_()
        # ---- a-3 + a-4 + a-5 (+ pose -> matrix): one launch ----
        rs = L.RaySetup()
        rs.mode, rs.n_batches = 0, len(batches)
        R = 0
        if not 1 <= len(batches) <= 2:
            raise ValueError("MappingStep: 1 or 2 keyframe batches (Mapper.py:379-393)")
        f32, i64 = torch.float32, torch.int64
        for bi, (b_, (c2ws, depths, colors, dirs_cam, indices, n, frame_base)) in enumerate(zip(rs.batch, batches)):
            Kb, P = depths.shape
            if joint and (frame_base < 0 or frame_base + Kb > K):
                raise ValueError(f"MappingStep: batch {bi} covers frames [{frame_base}, {frame_base + Kb}) of a {K}-frame window")
            if indices.numel() != Kb * n:
                raise ValueError(f"MappingStep: batch {bi} has {indices.numel()} indices, expected K*n = {Kb * n}")
            b_.c2ws = None if joint else cptr(c2ws, f32, Kb * 16, "c2ws")
            b_.depths, b_.colors = cptr(depths, f32, Kb * P, "depths"), cptr(colors, f32, Kb * P * 3, "colors")
            b_.dirs_cam, b_.indices = cptr(dirs_cam, f32, Kb * P * 3, "dirs_cam"), cptr(indices, i64, Kb * n, "indices")
            b_.P, b_.K, b_.n, b_.frame_base = P, Kb, n, frame_base
            R += Kb * n
        R_global, ray_off = R, 0
        if ray_range is not None:
            ray_off, ray_end = int(ray_range[0]), int(ray_range[1])
            if not 0 <= ray_off <= ray_end <= R_global:
                raise ValueError(f"MappingStep: ray_range {ray_range} outside the batch of {R_global} rays")
            R = ray_end - ray_off
        if R > self.max_rays:
            raise RuntimeError(f"MappingStep: {R} rays requested, buffers sized for max_rays={self.max_rays}")
        if self.perturb:
            cptr(t_rand, f32, R_global * S, "t_rand (R,S)")
        if has_holes:
            if u_pdf is None or (self.perturb and t_rand_uni is None):
                raise ValueError("MappingStep: rays without sensor depth need the draws of the no-depth branch "
                                 "(t_rand_uni (R,n_stratified), u_pdf (R,n_importance); Renderer.py:103-130); "
                                 "pass has_holes=False only if every sampled pixel has depth > 0")
            cptr(u_pdf, f32, R_global * self.zs.n_importance, "u_pdf (R,n_importance)")
            if self.perturb:
                cptr(t_rand_uni, f32, R_global * self.zs.n_stratified, "t_rand_uni (R,n_stratified)")
            if ray_off:                                     # rows of this rank's slice (contiguous)
                u_pdf = u_pdf[ray_off:ray_off + R]
                t_rand_uni = t_rand_uni[ray_off:ray_off + R] if t_rand_uni is not None else None
        if joint:
            cptr(cam_poses, f32, (K - 1) * 7, "cam_poses (K-1,7)"); cptr(c2w_fixed, f32, 12, "c2w_fixed")
        self.n_rays = R
        v = lambda t: ptr(t[:R]) if t is not None else None
        rs.cam_poses = ptr(cam_poses) if joint else None
        rs.c2w_fixed = ptr(c2w_fixed) if joint else None
        rs.bound, rs.require_depth, rs.zs = fs.meta.bound, 0, self.zs.args
        rs.t_rand = ptr(t_rand) if self.perturb else None
        rs.n_rays, rs.ray_offset = R, ray_off
        rs.rays_o, rs.rays_d, rs.gt_depth, rs.gt_color, rs.dirs_out = v(self.rays_o), v(self.rays_d), v(self.gt_depth), v(self.gt_color), v(self.dirs)
        rs.frame_id, rs.valid, rs.z = v(self.frame_id), v(self.valid), v(self.z)
        self._call("usl_ray_setup", byref(rs), st)
        if has_holes:                                      # a-6
            self._call("usl_zsample_nodepth", byref(self.zs.args), byref(fs.field), ptr(fs.beta), v(self.rays_o), v(self.rays_d), v(self.gt_depth),
                       v(self.valid), ptr(t_rand_uni) if self.perturb else None, ptr(u_pdf), None, R, v(self.z),
                       ptr(self.pdf_inds) if self.pdf_inds is not None else None, st)
        # ---- a-7, a-1, a-2: field query; a-8: compositing ----
        pts = L.Points()
        pts.x = None; pts.rays_o, pts.rays_d, pts.z, pts.valid = v(self.rays_o), v(self.rays_d), v(self.z), v(self.valid)
        pts.S, pts.n = S, R * S
        self._call("usl_field_fwd", byref(fs.field), byref(pts), v(self.raw), ptr(self.feat), ptr(self.jac) if joint else None, st)
        if fork:
            cur.wait_stream(side)
        # ---- a-8 + a-9 phase 1: compositing and the loss sums / counts in one launch (the mapper's mask is per ray) ----
        self._call("usl_composite_loss_fwd", byref(self.loss_args), v(self.raw), v(self.z), ptr(fs.beta), v(self.valid), R, S, v(self.gt_depth),
                   v(self.gt_color), v(self.term), v(self.punc), v(self.depth), v(self.rgb), v(self.dunc), ptr(self.acc), v(self.mask), st)
        if self.acc_hook is not None:
            self.acc_hook(self.acc)
        # ---- backward: loss gradient + compositing adjoint (+ loss value) in one launch, then the field ----
        self._call("usl_composite_loss_bwd", byref(self.loss_args), v(self.raw), v(self.z), ptr(fs.beta), v(self.valid), v(self.mask), R, S,
                   v(self.gt_depth), v(self.gt_color), v(self.depth), v(self.rgb), ptr(self.acc), None, ptr(self.jac) if joint else None,
                   byref(fs.meta.bound), v(self.d_raw), ptr(fs.g_beta), v(self.d_rays_o) if joint else None,
                   v(self.d_rays_d) if joint else None, ptr(self.loss), st)
        def pose_grads(st_):
            self._call("usl_pose_reduce", v(self.d_rays_o), v(self.d_rays_d), v(self.dirs), v(self.frame_id), v(self.valid), R, K, ptr(self.d_c2w), st_)
            call("usl_pose_matrix_bwd", ptr(cam_poses), ptr(self.d_c2w[1:K]), K - 1, ptr(self.d_pose[:K - 1]), st_)

        if joint and fork:                                 # needs only the ray gradients: runs beside the table scatter
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                pose_grads(side.cuda_stream)
        if self.rgb_grads_hook is None:
            self._call("usl_field_bwd", byref(fs.field), byref(pts), v(self.raw), ptr(self.feat), v(self.d_raw), ptr(fs.g_sdf_table),
                       ptr(fs.g_rgb_table), fs.g_mlp, ptr(fs.scratch), 3, st)
        else:
            # multi-GPU: the colour half (87 % of the gradient bytes) first, hand it to the collective, then the sdf half
            self._call("usl_field_bwd", byref(fs.field), byref(pts), v(self.raw), ptr(self.feat), v(self.d_raw), ptr(fs.g_sdf_table),
                       ptr(fs.g_rgb_table), fs.g_mlp, ptr(fs.scratch), 2, st)
            self.rgb_grads_hook(fs.g_rgb_table)
            self._call("usl_field_bwd", byref(fs.field), byref(pts), v(self.raw), ptr(self.feat), v(self.d_raw), ptr(fs.g_sdf_table),
                       ptr(fs.g_rgb_table), fs.g_mlp, ptr(fs.scratch), 1 | (4 if self.bwd_leave_room else 0), st)
        if joint:
            if fork:
                cur.wait_stream(side)
            else:
                pose_grads(st)
        return self.loss


class TrackingStep(_Profiled):
    """One tracking iteration: pose -> rays -> render -> median mask -> loss -> d loss / d (quat, trans).
    The hash tables and decoders are read-only here (the reference's dead table scatter is skipped)."""

    def __init__(self, meta: ops.FieldMeta, sdf_table, rgb_table, dec, beta, *, n_stratified, n_importance, truncation,
                 H, W, fx, fy, cx, cy, ignore_edge_h, ignore_edge_w, n_rays: int,
                 weights=(10.0, 200.0, 50.0, 1.0, 5.0), perturb: bool = True, mask_mode: str = "original"):
        dev = sdf_table.device
        self.fs = _FieldState(meta, sdf_table, rgb_table, dec, beta, with_grads=False)
        self.zs = ops.ZSampler(n_stratified, n_importance, truncation, dev)
        self.S = self.zs.S
        self.perturb = perturb
        self.cam = (H, W, float(fx), float(fy), float(cx), float(cy))
        self.win = (ignore_edge_h, H - ignore_edge_h, ignore_edge_w, W - ignore_edge_w)
        self.loss_args = ops.make_loss_args(truncation, weights[0], weights[1], weights[2], weights[3], weights[4],
                                            _mask_mode(mask_mode, 1))                       # cfg['t_mask_mode']
        R, S = n_rays, self.S
        self.R = R
        f32 = dict(device=dev, dtype=torch.float32)
        self.rays_o = torch.empty((R, 3), **f32); self.rays_d = torch.empty((R, 3), **f32)
        self.gt_depth = torch.empty((R,), **f32); self.gt_color = torch.empty((R, 3), **f32); self.dirs = torch.empty((R, 3), **f32)
        self.valid = torch.empty((R,), device=dev, dtype=torch.uint8); self.mask = torch.empty((R,), device=dev, dtype=torch.uint8)
        self.z = torch.zeros((R, S), **f32)
        self.raw = torch.empty((R, S, 4), **f32); self.jac = torch.empty((12 * R * S,), **f32)        # component-major [12][n] of THIS call's n = R*S
        self.term = torch.empty((R,), **f32); self.punc = torch.empty((R,), **f32); self.depth = torch.empty((R,), **f32)
        self.rgb = torch.empty((R, 3), **f32); self.dunc = torch.empty((R,), **f32)
        self._small = torch.zeros((L.LOSS_SLOTS + 12,), **f32)          # loss accumulators + d c2w: one memset per iteration
        self.acc = self._small[:L.LOSS_SLOTS]; self.d_c2w = self._small[L.LOSS_SLOTS:].view(1, 12)
        self.loss = torch.zeros((1,), **f32)
        self.median = torch.zeros((1,), **f32); self.ws = torch.empty((R,), **f32)
        self.d_raw = torch.empty((R, S, 4), **f32)
        self.d_rays_o = torch.empty((R, 3), **f32); self.d_rays_d = torch.empty((R, 3), **f32)
        self.d_pose = torch.zeros((1, 7), **f32)
        self._init_prof()

    def run(self, cam_pose, depth_img, color_img, indices, t_rand, best_loss=None, best_pose=None):
        """cam_pose (1,7) = [quat(real first), trans] (common.py:196-208); depth_img (H,W); color_img (H,W,3);
        indices: torch.randint(window_pixels, (R,)) (common.py:116); t_rand (R,S). Returns loss (1,);
        d loss/d cam_pose in self.d_pose; mean pixel uncertainty = acc[11]/acc[9].
        best_loss (1,) / best_pose (1,7), when given, are updated on the device the way Tracker.py:346-348 keeps its
        candidate pose (before the optimiser moves cam_pose)."""
        st = stream()
        fs, S, R = self.fs, self.S, self.R
        fs.refresh()
        H, W, fx, fy, cx, cy = self.cam
        H0, H1, W0, W1 = self.win
        f32 = torch.float32
        cptr(cam_pose, f32, 7, "cam_pose (1,7)"); cptr(depth_img, f32, H * W, "depth_img (H,W)"); cptr(color_img, f32, H * W * 3, "color_img (H,W,3)")
        cptr(indices, torch.int64, R, "indices (n_rays,)")
        if self.perturb:
            cptr(t_rand, f32, R * S, "t_rand (R,S)")
        self._small.zero_()
        rs = L.RaySetup()                                   # pose -> matrix, a-3, a-4, a-5 in one launch
        rs.mode, rs.n_batches = 1, 0
        rs.depth_img, rs.color_img, rs.win_indices = ptr(depth_img), ptr(color_img), ptr(indices)
        rs.H, rs.W, rs.H0, rs.H1, rs.W0, rs.W1 = H, W, H0, H1, W0, W1
        rs.fx, rs.fy, rs.cx, rs.cy = fx, fy, cx, cy
        rs.cam_poses = ptr(cam_pose)
        rs.bound, rs.require_depth, rs.zs = fs.meta.bound, 1, self.zs.args
        rs.t_rand = ptr(t_rand) if self.perturb else None
        rs.n_rays = R
        rs.rays_o, rs.rays_d, rs.gt_depth, rs.gt_color, rs.dirs_out = ptr(self.rays_o), ptr(self.rays_d), ptr(self.gt_depth), ptr(self.gt_color), ptr(self.dirs)
        rs.frame_id, rs.valid, rs.z = None, ptr(self.valid), ptr(self.z)
        self._call("usl_ray_setup", byref(rs), st)
        pts = L.Points()
        pts.x = None; pts.rays_o, pts.rays_d, pts.z, pts.valid = ptr(self.rays_o), ptr(self.rays_d), ptr(self.z), ptr(self.valid)
        pts.S, pts.n = S, R * S
        self._call("usl_field_fwd", byref(fs.field), byref(pts), ptr(self.raw), None, ptr(self.jac), st)
        self._call("usl_composite_fwd", ptr(self.raw), ptr(self.z), ptr(fs.beta), ptr(self.valid), R, S, ptr(self.term), ptr(self.punc),
                   ptr(self.depth), ptr(self.rgb), ptr(self.dunc), None, st)
        self._call("usl_depth_error_median", ptr(self.gt_depth), ptr(self.depth), ptr(self.valid), R, ptr(self.ws), ptr(self.median), st)
        self._call("usl_loss_fwd", byref(self.loss_args), ptr(self.raw), ptr(self.z), ptr(self.gt_depth), ptr(self.gt_color), ptr(self.valid),
                   ptr(self.punc), ptr(self.depth), ptr(self.rgb), ptr(self.median), R, S, ptr(self.acc), ptr(self.mask), st)
        self._call("usl_composite_loss_bwd", byref(self.loss_args), ptr(self.raw), ptr(self.z), ptr(fs.beta), ptr(self.valid), ptr(self.mask), R, S,
                   ptr(self.gt_depth), ptr(self.gt_color), ptr(self.depth), ptr(self.rgb), ptr(self.acc), None, ptr(self.jac),
                   byref(fs.meta.bound), ptr(self.d_raw), None, ptr(self.d_rays_o), ptr(self.d_rays_d), ptr(self.loss), st)
        self._call("usl_pose_reduce", ptr(self.d_rays_o), ptr(self.d_rays_d), ptr(self.dirs), None, ptr(self.valid), R, 1, ptr(self.d_c2w), st)
        call("usl_pose_matrix_bwd", ptr(cam_pose), ptr(self.d_c2w), 1, ptr(self.d_pose), st)
        if best_loss is not None:
            call("usl_track_keep_best", ptr(self.loss), ptr(cam_pose), ptr(best_loss), ptr(best_pose), st)
        return self.loss


class RenderImageStep(_Profiled):
    """Forward-only rendering of whole frames (Renderer.render_img, src/utils/Renderer.py:160-223; used for the
    visualisations and the end-of-run eval_rendering): get_rays -> depth-guided / no-depth z-sampling -> field query
    (no activation stash, no Jacobian) -> compositing, in chunks of ``chunk_rays`` consecutive pixels on preallocated
    buffers.  A pixel range [pixel_begin, pixel_end) can be rendered instead of the whole frame: that is the unit the
    multi-GPU path shards (rows of the image, no collective on the data path).

    RNG contract as for the other fused steps: draws are indexed by ray slot of the rendered range
    (t_rand (n,S), t_rand_uni (n,n_stratified), u_pdf (n,n_importance)); the drop-in modules.Renderer.render_img keeps
    the reference's per-chunk compacted draw order instead."""

    def __init__(self, meta: ops.FieldMeta, sdf_table, rgb_table, dec, beta, *, n_stratified, n_importance, truncation,
                 H, W, fx, fy, cx, cy, chunk_rays: int = 131072, perturb: bool = True):
        dev = sdf_table.device
        self.fs = _FieldState(meta, sdf_table, rgb_table, dec, beta, with_grads=False)
        self.zs = ops.ZSampler(n_stratified, n_importance, truncation, dev)
        self.S = self.zs.S
        self.perturb = perturb
        self.cam = (int(H), int(W), float(fx), float(fy), float(cx), float(cy))
        C = self.chunk = int(min(chunk_rays, H * W))
        self.sample_major = True        # set False for ray-major point order (ablation: 13.2 vs 10.7 ms per frame)
        f32 = dict(device=dev, dtype=torch.float32)
        self.rays_o = torch.empty((C, 3), **f32); self.rays_d = torch.empty((C, 3), **f32)
        self.gt_depth = torch.empty((C,), **f32); self.valid = torch.empty((C,), device=dev, dtype=torch.uint8)
        self.z = torch.zeros((C, self.S), **f32); self.raw = torch.empty((C, self.S, 4), **f32)
        self._init_prof()

    def alloc_outputs(self, n):
        dev = self.z.device
        f32 = dict(device=dev, dtype=torch.float32)
        return dict(depth=torch.empty((n,), **f32), color=torch.empty((n, 3), **f32), term=torch.empty((n,), **f32),
                    pixel_unc=torch.empty((n,), **f32), depth_unc=torch.empty((n,), **f32))

    def run(self, c2w, depth_img, t_rand, t_rand_uni=None, u_pdf=None, pixel_begin: int = 0, pixel_end: Optional[int] = None,
            out: Optional[dict] = None, has_holes: bool = True):
        """c2w (4,4); depth_img (H,W) sensor depth (0 = hole). Returns dict of fp32 tensors over the pixel range:
        depth, color (n,3), term, pixel_unc, depth_unc (the reference converts all but colour to float64 on return)."""
        st = stream()
        fs, S = self.fs, self.S
        fs.refresh()
        H, W, fx, fy, cx, cy = self.cam
        pixel_end = H * W if pixel_end is None else pixel_end
        n_total = pixel_end - pixel_begin
        if self.perturb:
            cptr(t_rand, torch.float32, n_total * S, "t_rand (n,S)")
        if has_holes:
            if u_pdf is None or (self.perturb and t_rand_uni is None):
                raise ValueError("RenderImageStep: pixels without sensor depth need t_rand_uni / u_pdf (Renderer.py:103-130); "
                                 "pass has_holes=False only for hole-free depth images")
            cptr(u_pdf, torch.float32, n_total * self.zs.n_importance, "u_pdf (n,n_importance)")
        out = out if out is not None else self.alloc_outputs(n_total)
        c2w = L.f32c(c2w); depth_img = L.f32c(depth_img)
        for c0 in range(0, n_total, self.chunk):
            n = min(self.chunk, n_total - c0)
            sl = lambda t: ptr(t[c0:c0 + n]) if t is not None else None
            rs = L.RaySetup()
            rs.mode, rs.n_batches = 2, 0
            rs.depth_img, rs.color_img, rs.win_indices = ptr(depth_img), None, None
            rs.H, rs.W = H, W
            rs.fx, rs.fy, rs.cx, rs.cy = fx, fy, cx, cy
            rs.c2w = ptr(c2w)
            rs.bound, rs.require_depth, rs.zs = fs.meta.bound, 0, self.zs.args
            rs.t_rand = sl(t_rand) if self.perturb else None
            rs.n_rays, rs.pixel_begin = n, pixel_begin + c0
            rs.rays_o, rs.rays_d, rs.gt_depth = ptr(self.rays_o), ptr(self.rays_d), ptr(self.gt_depth)
            rs.gt_color, rs.dirs_out, rs.frame_id = None, None, None
            rs.valid, rs.z = ptr(self.valid), ptr(self.z)
            self._call("usl_ray_setup", byref(rs), st)
            if has_holes:
                self._call("usl_zsample_nodepth", byref(self.zs.args), byref(fs.field), ptr(fs.beta), ptr(self.rays_o), ptr(self.rays_d),
                           ptr(self.gt_depth), None, sl(t_rand_uni) if self.perturb else None, sl(u_pdf), None, n, ptr(self.z), None, st)
            pts = L.Points()
            pts.x = None; pts.rays_o, pts.rays_d, pts.z, pts.valid = ptr(self.rays_o), ptr(self.rays_d), ptr(self.z), None
            pts.S, pts.n = S, n * S
            pts.sample_major = 1 if self.sample_major else 0   # consecutive rays = neighbouring pixels: a warp's points share cells
            self._call("usl_field_fwd", byref(fs.field), byref(pts), ptr(self.raw), None, None, st)
            self._call("usl_composite_fwd", ptr(self.raw), ptr(self.z), ptr(fs.beta), None, n, S, sl(out["term"]), sl(out["pixel_unc"]),
                       sl(out["depth"]), sl(out["color"]), sl(out["depth_unc"]), None, st)
        return out


class RenderMetrics:
    """eval_rendering's per-frame metrics (src/tools/eval_recon.py:276-299) over rendered frames where they lie on the device:
    PSNR over the pixels with sensor depth (mse_loss(gt_color[gt_depth > 0], color[gt_depth > 0]), psnr = -10 log10) and the
    depth L1 (mean |gt_depth - depth| over the same pixels), one usl_render_metrics launch per frame, sums in double, no host
    read before result().  MS-SSIM and LPIPS (third-party networks) are not part of it.

    add() takes RenderImageStep.run's outputs (out['color'], out['depth']) or modules.Renderer.render_img's (color, depth):
    any shape with 3 / 1 trailing channels, fp32 or the float64 render_img returns (widened fp32, cast back exactly); the
    dataset colour is cast to fp32 (the reference keeps it float64, datasets.py:87-91: a difference below 6e-8 per value)."""

    def __init__(self, device, max_frames: int = 4096):
        self.acc = torch.zeros((max_frames, 3), device=device, dtype=torch.float64)
        self.n = 0

    def add(self, color: torch.Tensor, depth: torch.Tensor, gt_color: torch.Tensor, gt_depth: torch.Tensor):
        if self.n >= self.acc.shape[0]:
            raise RuntimeError(f"RenderMetrics: sized for {self.acc.shape[0]} frames")
        n = gt_depth.numel()
        if depth.numel() != n or color.numel() != 3 * n or gt_color.numel() != 3 * n:
            raise ValueError("RenderMetrics.add: color / gt_color need 3 values and depth one value per pixel of gt_depth")
        c, d, gc, gd = [L.f32c(t).reshape(-1) for t in (color, depth, gt_color, gt_depth)]
        for t, nm in ((c, "color"), (d, "depth"), (gc, "gt_color"), (gd, "gt_depth")):
            cptr(t, torch.float32, None, nm)
        call("usl_render_metrics", ptr(gc), ptr(gd), ptr(c), ptr(d), n, ptr(self.acc[self.n]), stream())
        self.n += 1

    @staticmethod
    def finalize(acc: torch.Tensor) -> dict:
        """acc (F,3) float64 rows [sum sq. colour error, sum |depth error|, pixel count] -> the dict eval_rendering writes (its
        third-party entries aside) + the per-frame values.  A frame without a single pixel of depth gives NaN, as the
        reference's mean over an empty selection does."""
        a = acc.detach().to("cpu", torch.float64)
        mse = a[:, 0] / (3.0 * a[:, 2])
        psnr = -10.0 * torch.log10(mse)
        l1 = a[:, 1] / a[:, 2]
        return {"avg_psnr": float(psnr.mean()) if len(a) else float("nan"), "depth_l1_render": float(l1.mean()) if len(a) else float("nan"),
                "psnr": psnr, "depth_l1": l1, "mse": mse, "frames": int(len(a))}

    def result(self) -> dict:
        return self.finalize(self.acc[:self.n])            # the one host read of the sequence


class DenseSdfQuery:
    """Mesher.get_grid_uniform + eval_points (SDF channel) over a y-slab of the 1 cm query grid
    (src/utils/Mesher.py:134-195,219-227); points are generated in-kernel from the per-axis coordinates."""

    def __init__(self, meta: ops.FieldMeta, sdf_table, rgb_table, dec, axes: Sequence[torch.Tensor]):
        self.meta = meta
        self.field = meta.pack(sdf_table, rgb_table, list(dec))
        self._keep = (sdf_table, rgb_table, list(dec))
        self.ax, self.ay, self.az = [a.contiguous().float() for a in axes]
        self.nx, self.ny, self.nz = self.ax.numel(), self.ay.numel(), self.az.numel()

    def slab_points(self, y_begin, y_end):
        return (y_end - y_begin) * self.nx * self.nz

    def run(self, y_begin: int, y_end: int, out: Optional[torch.Tensor] = None):
        n = self.slab_points(y_begin, y_end)
        self.field = self.meta.pack(self._keep[0], self._keep[1], self._keep[2])    # pick up rebound tensors
        if out is None:
            out = torch.empty((n,), device=self.ax.device, dtype=torch.float32)
        call("usl_sdf_query_grid", byref(self.field), ptr(self.ax), ptr(self.ay), ptr(self.az), self.nx, self.ny, self.nz,
             y_begin, y_end, ptr(out), stream())
        return out
