"""Thin torch layer over the C-ABI: tensor allocation, struct packing and autograd plumbing only.

Every function here forwards to one `usl_*` entry point of lib/libunislam_b200.so; the arithmetic
lives in csrc/*.cu.  Names and argument meaning mirror the reference host code they serve
(src/common.py, src/utils/Renderer.py, src/networks/decoders.py, src/Mapper.py, src/Tracker.py).
"""
from ctypes import byref
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib as L
from ._lib import call, f32c, ptr, stream

N_LEVELS, N_FEATS, C_DIM, HIDDEN = 16, 2, 32, 16


def stash_floats(n_points: int) -> int:
    """Size of the activation stash usl_field_fwd writes for usl_field_bwd (features, hidden pre-activations, clamped coordinates)."""
    from ctypes import c_int64
    out = c_int64(0)
    call("usl_field_stash_floats", int(n_points), byref(out))
    return int(out.value)


# ------------------------------------------------------------------------------------------------
# B1: tinycudann.Encoding
# ------------------------------------------------------------------------------------------------
class _GridEncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, params, grid):
        x = f32c(x)
        n = x.shape[0]
        y = torch.empty((n, grid.n_levels * N_FEATS), device=x.device, dtype=torch.float32)
        call("usl_grid_encode_fwd", byref(grid), ptr(params), ptr(x), n, ptr(y), stream())
        ctx.grid = grid
        ctx.save_for_backward(x, params)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, params = ctx.saved_tensors
        grid = ctx.grid
        dy = f32c(dy)
        n = x.shape[0]
        dx = dparams = None
        if ctx.needs_input_grad[1]:
            dparams = torch.zeros_like(params)
            call("usl_grid_encode_bwd_params", byref(grid), ptr(x), ptr(dy), n, ptr(dparams), stream())
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            call("usl_grid_encode_bwd_input", byref(grid), ptr(params), ptr(x), ptr(dy), n, ptr(dx), stream())
        return dx, dparams, None


def grid_encode(x: torch.Tensor, params: torch.Tensor, grid: L.Grid) -> torch.Tensor:
    """y (N, 2L) = HashGrid(x (N,3)); differentiable wrt x and params (decoders.py:103)."""
    return _GridEncodeFn.apply(x, params, grid)


def grid_corner_indices(x: torch.Tensor, grid: L.Grid) -> torch.Tensor:
    """(N, L, 8) within-level entry indices of every corner (parity inspection)."""
    x = f32c(x)
    out = torch.empty((x.shape[0], grid.n_levels, 8), device=x.device, dtype=torch.int32)
    call("usl_grid_corner_indices", byref(grid), ptr(x), x.shape[0], ptr(out), stream())
    return out.to(torch.int64) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------
# decoder weight packing
# ------------------------------------------------------------------------------------------------
@dataclass
class DecoderLayout:
    """Where one decoder's tensors live. Variant 'A': nn.Linear stacks (decoders.py:72-84), 'B': tcnn.Network
    flat params restated in fp32 (decoders.py:49-70): params[0:512] = W1 (16,32), params[512:768] = Wout (16,16)."""
    variant: str
    n_out: int
    out_act: int

    def n_params_tensors(self) -> int:
        return 6 if self.variant == "A" else 1

    def pack(self, tensors: Sequence[torch.Tensor]) -> L.Mlp:
        m = L.Mlp()
        m.n_out, m.out_act = self.n_out, self.out_act
        if self.variant == "A":
            w1, b1, w2, b2, wo, bo = tensors
            m.w1, m.b1, m.w2, m.b2, m.wo, m.bo = ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(wo), ptr(bo)
            m.n_hidden = 2
        else:
            (p,) = tensors
            base = ptr(p)
            m.w1, m.wo = base, base + 512 * 4
            m.b1 = m.w2 = m.b2 = m.bo = None
            m.n_hidden = 1
        return m


SDF_ACT, RGB_ACT = L.ACT_TANH, L.ACT_SIGMOID


@dataclass
class FieldMeta:
    """Static description of scene_rep + Decoders (grids, decoder variant, bound)."""
    grid_sdf: L.Grid
    grid_rgb: L.Grid
    variant: str
    bound: L.Bound

    @property
    def layouts(self):
        return DecoderLayout(self.variant, 1, SDF_ACT), DecoderLayout(self.variant, 3, RGB_ACT)

    @property
    def n_dec_tensors(self) -> int:
        return self.layouts[0].n_params_tensors()

    def pack(self, sdf_table, rgb_table, dec_tensors: Sequence[torch.Tensor]) -> L.Field:
        f = L.Field()
        f.grid[0], f.grid[1] = self.grid_sdf, self.grid_rgb
        f.table[0], f.table[1] = ptr(sdf_table), ptr(rgb_table)
        k = self.n_dec_tensors
        ls, lr = self.layouts
        f.mlp[0] = ls.pack(dec_tensors[:k])
        f.mlp[1] = lr.pack(dec_tensors[k:2 * k])
        for d in range(3):
            f.bound_lo[d] = self.bound.lo[d]; f.bound_hi[d] = self.bound.hi[d]
        return f

    def pack_grads(self, grad_tensors: Sequence[torch.Tensor]):
        k = self.n_dec_tensors
        ls, lr = self.layouts
        arr = (L.Mlp * 2)()
        arr[0] = ls.pack(grad_tensors[:k]); arr[1] = lr.pack(grad_tensors[k:2 * k])
        return arr


def bwd_scratch_floats(field: L.Field) -> int:
    from ctypes import c_int64
    n = c_int64(0)
    call("usl_field_bwd_scratch_floats", byref(field), byref(n))
    return int(n.value)


def bwd_scratch(field: L.Field, device) -> torch.Tensor:
    """Zero-filled workspace for usl_field_bwd's replicated coarse levels."""
    return torch.zeros(max(bwd_scratch_floats(field), 4), device=device, dtype=torch.float32)


def _points_from_rays(rays_o, rays_d, z, valid=None) -> L.Points:
    p = L.Points()
    p.x = None
    p.rays_o, p.rays_d, p.z = ptr(rays_o), ptr(rays_d), ptr(z)
    p.valid = ptr(valid) if valid is not None else None
    p.S = z.shape[1]
    p.n = z.shape[0] * z.shape[1]
    return p


def _points_from_x(x) -> L.Points:
    p = L.Points()
    p.x = ptr(x)
    p.rays_o = p.rays_d = p.z = p.valid = None
    p.S = 1
    p.n = x.shape[0]
    return p


# ------------------------------------------------------------------------------------------------
# B3: Decoders.forward on explicit points (Mesher.eval_points, Decoders drop-in)
# ------------------------------------------------------------------------------------------------
class _FieldPointsFn(torch.autograd.Function):
    """raw (N,4) = decoders(p (N,3)); differentiable wrt tables, decoder weights and p."""

    @staticmethod
    def forward(ctx, meta: FieldMeta, x, sdf_table, rgb_table, *dec):
        x = f32c(x)
        n = x.shape[0]
        need_p = any(ctx.needs_input_grad[2:])
        need_x = ctx.needs_input_grad[1]
        raw = torch.empty((n, 4), device=x.device, dtype=torch.float32)
        feat = torch.empty((stash_floats(n),), device=x.device, dtype=torch.float32) if need_p else None
        jac = torch.empty((12, n), device=x.device, dtype=torch.float32) if need_x else None
        f = meta.pack(sdf_table, rgb_table, dec)
        pts = _points_from_x(x)
        call("usl_field_fwd", byref(f), byref(pts), ptr(raw), ptr(feat), ptr(jac), stream())
        ctx.meta = meta
        ctx.save_for_backward(x, sdf_table, rgb_table, raw, feat, jac, *dec)
        return raw

    @staticmethod
    def backward(ctx, d_raw):
        x, sdf_table, rgb_table, raw, feat, jac, *dec = ctx.saved_tensors
        meta = ctx.meta
        d_raw = f32c(d_raw)
        dx = None
        if jac is not None:
            dx = torch.einsum("no,odn->nd", d_raw, jac.view(4, 3, -1))
        gs = gr = None
        gdec = [None] * len(dec)
        if feat is not None:
            gs, gr = torch.zeros_like(sdf_table), torch.zeros_like(rgb_table)
            gdec = [torch.zeros_like(t) for t in dec]
            f = meta.pack(sdf_table, rgb_table, dec)
            pts = _points_from_x(x)
            call("usl_field_bwd", byref(f), byref(pts), ptr(raw), ptr(feat), ptr(d_raw), ptr(gs), ptr(gr),
                 meta.pack_grads(gdec), ptr(bwd_scratch(f, x.device)), 3, stream())
        return (None, dx, gs, gr, *gdec)


def field_points(meta: FieldMeta, x, sdf_table, rgb_table, dec: Sequence[torch.Tensor]) -> torch.Tensor:
    return _FieldPointsFn.apply(meta, x, sdf_table, rgb_table, *dec)


def field_sdf_points(meta: FieldMeta, x, sdf_table, rgb_table, dec) -> torch.Tensor:
    """Decoders.get_raw_sdf, forward only (Renderer.py:121)."""
    x = f32c(x)
    out = torch.empty((x.shape[0],), device=x.device, dtype=torch.float32)
    f = meta.pack(sdf_table, rgb_table, dec)
    pts = _points_from_x(x)
    call("usl_field_sdf", byref(f), byref(pts), ptr(out), stream())
    return out


# ------------------------------------------------------------------------------------------------
# B4: render (field query on ray samples + compositing), one autograd node
# ------------------------------------------------------------------------------------------------
class _RenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meta: FieldMeta, rays_o, rays_d, z_vals, beta, sdf_table, rgb_table, *dec):
        rays_o, rays_d, z_vals = f32c(rays_o), f32c(rays_d), f32c(z_vals)
        R, S = z_vals.shape
        dev = z_vals.device
        need_rays = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        need_p = any(ctx.needs_input_grad[5:])
        raw = torch.empty((R, S, 4), device=dev, dtype=torch.float32)
        feat = torch.empty((stash_floats(R * S),), device=dev, dtype=torch.float32) if need_p else None
        jac = torch.empty((12, R * S), device=dev, dtype=torch.float32) if need_rays else None
        f = meta.pack(sdf_table, rgb_table, dec)
        pts = _points_from_rays(rays_o, rays_d, z_vals)
        st = stream()
        call("usl_field_fwd", byref(f), byref(pts), ptr(raw), ptr(feat), ptr(jac), st)
        term = torch.empty((R,), device=dev); punc = torch.empty((R,), device=dev); depth = torch.empty((R,), device=dev)
        rgb = torch.empty((R, 3), device=dev); dunc = torch.empty((R,), device=dev)
        call("usl_composite_fwd", ptr(raw), ptr(z_vals), ptr(beta), None, R, S, ptr(term), ptr(punc), ptr(depth), ptr(rgb),
             ptr(dunc), None, st)
        ctx.meta = meta
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(rays_o, rays_d, z_vals, beta, sdf_table, rgb_table, raw, feat, jac, *dec)
        sdf = raw[..., 3]
        return term, punc, depth, rgb, sdf, dunc

    @staticmethod
    def backward(ctx, g_term, g_punc, g_depth, g_rgb, g_sdf, g_dunc):
        rays_o, rays_d, z_vals, beta, sdf_table, rgb_table, raw, feat, jac, *dec = ctx.saved_tensors
        meta = ctx.meta
        R, S = z_vals.shape
        dev = z_vals.device
        c = lambda t: None if t is None else f32c(t)
        g_term, g_punc, g_depth, g_rgb, g_sdf, g_dunc = map(c, (g_term, g_punc, g_depth, g_rgb, g_sdf, g_dunc))
        d_raw = torch.empty((R, S, 4), device=dev, dtype=torch.float32)
        d_beta = torch.zeros((1,), device=dev, dtype=torch.float32)
        d_o = torch.empty((R, 3), device=dev) if jac is not None else None
        d_d = torch.empty((R, 3), device=dev) if jac is not None else None
        st = stream()
        call("usl_composite_bwd", ptr(raw), ptr(z_vals), ptr(beta), None, R, S, ptr(g_term), ptr(g_punc), ptr(g_depth),
             ptr(g_rgb), ptr(g_dunc), ptr(g_sdf), ptr(jac), byref(meta.bound), ptr(d_raw), ptr(d_beta), ptr(d_o), ptr(d_d), st)
        gs = gr = None
        gdec = [None] * len(dec)
        if feat is not None:
            want_s, want_r = ctx.needs_input_grad[5], ctx.needs_input_grad[6]
            want_dec = any(ctx.needs_input_grad[7:])
            gs = torch.zeros_like(sdf_table) if want_s else None
            gr = torch.zeros_like(rgb_table) if want_r else None
            if want_dec:
                gdec = [torch.zeros_like(t) for t in dec]
            f = meta.pack(sdf_table, rgb_table, dec)
            pts = _points_from_rays(rays_o, rays_d, z_vals)
            call("usl_field_bwd", byref(f), byref(pts), ptr(raw), ptr(feat), ptr(d_raw), ptr(gs), ptr(gr),
                 meta.pack_grads(gdec) if want_dec else None, ptr(bwd_scratch(f, dev)), 3, st)
        return (None, d_o if ctx.needs_input_grad[1] else None, d_d if ctx.needs_input_grad[2] else None, None,
                d_beta if ctx.needs_input_grad[4] else None, gs, gr, *gdec)


def render_rays(meta: FieldMeta, rays_o, rays_d, z_vals, beta, sdf_table, rgb_table, dec: Sequence[torch.Tensor]):
    """(term, pixel_unc, depth, rgb, sdf, depth_unc) for R rays x S samples (Renderer.py:132-152)."""
    return _RenderFn.apply(meta, rays_o, rays_d, z_vals, beta, sdf_table, rgb_table, *dec)


# ------------------------------------------------------------------------------------------------
# sampling helpers (forward only; RNG draws come from torch exactly as in the reference)
# ------------------------------------------------------------------------------------------------
class ZSampler:
    """Holds the torch.linspace tables and constants of Renderer.render_batch_ray (Renderer.py:77-95)."""

    def __init__(self, n_stratified: int, n_importance: int, truncation: float, device):
        self.n_stratified, self.n_importance = n_stratified, n_importance
        self.t_uni = torch.linspace(0., 1., steps=n_stratified, device=device)
        self.t_surf = torch.linspace(0., 1., steps=n_importance, device=device)
        a = L.ZSampleArgs()
        a.n_stratified, a.n_importance = n_stratified, n_importance
        a.c_surf_lo = 1.5 * truncation          # python double product, then fp32 (Renderer.py:92)
        a.c_surf_span = 3 * truncation
        a.t_uni, a.t_surf = ptr(self.t_uni), ptr(self.t_surf)
        self.args = a

    @property
    def S(self):
        return self.n_stratified + self.n_importance

    def depth_guided(self, gt_depth, z, t_rand=None, valid=None, row_map=None):
        call("usl_zsample_depth", byref(self.args), ptr(gt_depth), ptr(valid), ptr(t_rand), ptr(row_map),
             gt_depth.shape[0], ptr(z), stream())

    def no_depth(self, field: L.Field, beta, rays_o, rays_d, gt_depth, z, u_pdf, t_rand_uni=None, valid=None,
                 row_map=None, pdf_inds=None):
        call("usl_zsample_nodepth", byref(self.args), byref(field), ptr(beta), ptr(rays_o), ptr(rays_d), ptr(gt_depth),
             ptr(valid), ptr(t_rand_uni), ptr(u_pdf), ptr(row_map), gt_depth.shape[0], ptr(z), ptr(pdf_inds), stream())


def sample_keyframe_rays(c2ws, depths, colors, dirs_cam, indices, n, frame_base=0, out=None):
    """get_samples_all (common.py:152-166). Returns rays_o, rays_d, gt_depth, gt_color, dirs, frame_id."""
    K, P = depths.shape
    M = K * n
    dev = depths.device
    if out is None:
        out = (torch.empty((M, 3), device=dev), torch.empty((M, 3), device=dev), torch.empty((M,), device=dev),
               torch.empty((M, 3), device=dev), torch.empty((M, 3), device=dev), torch.empty((M,), device=dev, dtype=torch.int32))
    ro, rd, gd, gc, dirs, fid = out
    call("usl_sample_keyframe_rays", ptr(f32c(c2ws)), ptr(depths), ptr(colors), ptr(dirs_cam), ptr(indices), K, P, n,
         frame_base, ptr(ro), ptr(rd), ptr(gd), ptr(gc), ptr(dirs), ptr(fid), stream())
    return out


def sample_window_rays(c2w, depth, color, H0, H1, W0, W1, fx, fy, cx, cy, indices, out=None):
    """get_samples for one frame (common.py:168-180)."""
    H, W = depth.shape[-2:]
    n = indices.shape[0]
    dev = depth.device
    if out is None:
        out = (torch.empty((n, 3), device=dev), torch.empty((n, 3), device=dev), torch.empty((n,), device=dev),
               torch.empty((n, 3), device=dev), torch.empty((n, 3), device=dev))
    ro, rd, gd, gc, dirs = out
    call("usl_sample_window_rays", ptr(f32c(c2w)), ptr(depth), ptr(color), H, W, H0, H1, W0, W1, fx, fy, cx, cy,
         ptr(indices), n, ptr(ro), ptr(rd), ptr(gd), ptr(gc), ptr(dirs), stream())
    return out


def image_rays(c2w, H, W, fx, fy, cx, cy):
    """get_rays (common.py:210-228)."""
    dev = c2w.device
    ro, rd = torch.empty((H, W, 3), device=dev), torch.empty((H, W, 3), device=dev)
    call("usl_image_rays", ptr(f32c(c2w)), H, W, fx, fy, cx, cy, ptr(ro), ptr(rd), stream())
    return ro, rd


def bbox_prefilter(rays_o, rays_d, gt_depth, bound: L.Bound, require_depth: bool, want_t=False):
    n = rays_o.shape[0]
    valid = torch.empty((n,), device=rays_o.device, dtype=torch.uint8)
    t = torch.empty((n,), device=rays_o.device) if want_t else None
    call("usl_bbox_prefilter", ptr(rays_o), ptr(rays_d), ptr(gt_depth), n, byref(bound), int(require_depth), ptr(t), ptr(valid), stream())
    return (valid, t) if want_t else valid


def make_loss_args(truncation, w_fs, w_center, w_tail, w_depth, w_color, mode) -> L.LossArgs:
    a = L.LossArgs()
    a.truncation = truncation
    a.truncation_center = 0.4 * truncation       # python double product, then fp32 (Mapper.py:160-161)
    a.w_sdf_fs, a.w_sdf_center, a.w_sdf_tail, a.w_depth, a.w_color = w_fs, w_center, w_tail, w_depth, w_color
    a.mode = mode
    return a


def pose_to_matrix(pose):
    K = pose.shape[0]
    out = torch.empty((K, 4, 4), device=pose.device)
    call("usl_pose_to_matrix", ptr(f32c(pose)), K, ptr(out), stream())
    return out
