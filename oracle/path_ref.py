"""oracle/path_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (PyTorch) restatement of the Uni-SLAM per-frame differentiable-rendering path:
ray generation, depth-guided / importance z-sampling, hash-grid + MLP field query,
SDF->weight compositing, masked losses and the parameter / pose gradients (via autograd).

Every function cites the reference lines it restates.  RNG draws are *arguments* (the
reference draws them with torch.randint / torch.rand; tests record the reference's draws and
feed the same tensors to the oracle and to the CUDA path).

Pinned here against the UNMODIFIED reference Python code (run through oracle/shims) by
oracle/gen_golden.py -> tests/golden/*.npz.  The tcnn grid arithmetic underneath is
"parity unpinned" (see grid_ref.py).
"""
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import grid_ref


# ----------------------------------------------------------------------------------------------
# pose parametrisation  (src/common.py:182-208, pytorch3d.transforms @47d5dc88)
# ----------------------------------------------------------------------------------------------
def quaternion_to_matrix(q: torch.Tensor) -> torch.Tensor:
    """pytorch3d.transforms.quaternion_to_matrix: real-first, un-normalised input."""
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def _sqrt_positive_part(x: torch.Tensor) -> torch.Tensor:
    ret = torch.zeros_like(x)
    m = x > 0
    ret[m] = torch.sqrt(x[m])
    return ret


def matrix_to_quaternion(matrix: torch.Tensor) -> torch.Tensor:
    """pytorch3d.transforms.matrix_to_quaternion (4-candidate form, sign standardised w>=0)."""
    batch_dim = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(matrix.reshape(batch_dim + (9,)), dim=-1)
    q_abs = _sqrt_positive_part(torch.stack([
        1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
        1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1))
    quat_by_rijk = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    flr = torch.tensor(0.1).to(dtype=q_abs.dtype, device=q_abs.device)
    quat_candidates = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    out = quat_candidates[F.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5, :].reshape(batch_dim + (4,))
    return torch.where(out[..., 0:1] < 0, -out, out)


def cam_pose_to_matrix(batch_poses: torch.Tensor) -> torch.Tensor:
    """src/common.py:196-208. pose = [qw,qx,qy,qz,tx,ty,tz]."""
    c2w = torch.eye(4, device=batch_poses.device, dtype=batch_poses.dtype).unsqueeze(0).repeat(batch_poses.shape[0], 1, 1)
    c2w[:, :3, :3] = quaternion_to_matrix(batch_poses[:, :4])
    c2w[:, :3, 3] = batch_poses[:, 4:]
    return c2w


def matrix_to_cam_pose(batch_matrices: torch.Tensor) -> torch.Tensor:
    """src/common.py:182-194 (RT=True)."""
    return torch.cat([matrix_to_quaternion(batch_matrices[:, :3, :3]), batch_matrices[:, :3, 3]], dim=-1)


# ----------------------------------------------------------------------------------------------
# ray generation  (src/common.py:35-46, 95-180, 210-228)
# ----------------------------------------------------------------------------------------------
def camera_dirs(H, W, fx, fy, cx, cy) -> torch.Tensor:
    """get_camera_rays (common.py:35-46), OpenGL: ((i-cx)/fx, -(j-cy)/fy, -1), shape (H,W,3)."""
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32), torch.arange(H, dtype=torch.float32), indexing="xy")
    return torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)


def rotate_dirs(dirs: torch.Tensor, c2ws: torch.Tensor):
    """common.py:102-105 / 160-162: rays_d = sum(dirs[...,None,:] * R, -1); rays_o = t.
    dirs (B,n,3), c2ws (B,4,4) -> (B,n,3) x2."""
    rays_d = torch.sum(dirs.unsqueeze(-2) * c2ws[:, None, :3, :3], -1)
    rays_o = c2ws[:, None, :3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def sample_tracking_rays(H0, H1, W0, W1, n, fx, fy, cx, cy, c2ws, depths, colors, indices):
    """get_samples -> get_sample_uv -> select_uv -> get_rays_from_uv (common.py:95-150,168-180).
    indices: the torch.randint((H1-H0)*(W1-W0), (n*b,)) draw of select_uv (common.py:116)."""
    b = c2ws.shape[0]
    Wc = W1 - W0
    if not (H0 == 0 and W0 == 0):
        depths = depths[:, H0:H1, W0:W1]
        colors = colors[:, H0:H1, W0:W1]
    # meshgrid+transpose of linspace(W0,W1-1) x linspace(H0,H1-1) flattened row-major (common.py:144-148)
    i = (indices % Wc).to(torch.float32) + float(W0)
    j = torch.div(indices, Wc, rounding_mode="floor").to(torch.float32) + float(H0)
    ind = indices.reshape(b, -1)
    i = i.reshape(b, -1); j = j.reshape(b, -1)
    d = torch.gather(depths.reshape(b, -1), 1, ind)
    c = torch.gather(colors.reshape(b, -1, 3), 1, ind.unsqueeze(-1).expand(-1, -1, 3))
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)
    rays_o, rays_d = rotate_dirs(dirs, c2ws)
    return rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), d.reshape(-1), c.reshape(-1, 3)


def sample_mapping_rays(c2ws, depths, colors, rays_d_cam, indices):
    """get_samples_all (common.py:152-166). depths (K,P), colors (K,P,3), rays_d_cam (K,P,3),
    indices: torch.randint(P,(n*K,)) draw (common.py:155)."""
    b = c2ws.shape[0]
    ind = indices.reshape(b, -1)
    sample_depth = torch.gather(depths, 1, ind)
    sample_color = torch.gather(colors, 1, ind.unsqueeze(-1).expand(-1, -1, 3))
    rays_o, rays_d = rotate_dirs(rays_d_cam, c2ws)
    rays_d = torch.gather(rays_d, 1, ind.unsqueeze(-1).expand(-1, -1, 3))
    rays_o = torch.gather(rays_o, 1, ind.unsqueeze(-1).expand(-1, -1, 3))
    return rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), sample_depth.reshape(-1), sample_color.reshape(-1, 3)


def full_image_rays(H, W, fx, fy, cx, cy, c2w):
    """get_rays (common.py:210-228)."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i = i.t(); j = j.t()
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1).reshape(H, W, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def bbox_exit(rays_o, rays_d, bound):
    """Mapper.py:396-401 / Tracker.py:177-183 / Renderer.py:106-111: t_exit = min_d max(t_lo,t_hi)."""
    t = (bound.unsqueeze(0) - rays_o.detach().unsqueeze(-1)) / rays_d.detach().unsqueeze(-1)
    t, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
    return t


# ----------------------------------------------------------------------------------------------
# z sampling  (src/utils/Renderer.py:42-57, 77-130; src/common.py:49-85)
# ----------------------------------------------------------------------------------------------
def perturb(z_vals, t_rand):
    """Renderer.perturbation (Renderer.py:42-57) with the torch.rand draw passed in."""
    mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
    upper = torch.cat([mids, z_vals[..., -1:]], -1)
    lower = torch.cat([z_vals[..., :1], mids], -1)
    return lower + (upper - lower) * t_rand


def zvals_with_depth(gt_nonzero, n_stratified, n_importance, truncation, t_rand=None):
    """Renderer.py:81-100. gt_nonzero (R,1) > 0. t_rand (R,S) or None (perturb off)."""
    dev = gt_nonzero.device
    t_uni = torch.linspace(0., 1., steps=n_stratified, device=dev)
    t_surf = torch.linspace(0., 1., steps=n_importance, device=dev)
    z_surf = gt_nonzero.expand(-1, n_importance) - (1.5 * truncation) + (3 * truncation * t_surf)
    z_free = 0.0 + 1.2 * gt_nonzero.expand(-1, n_stratified) * t_uni
    z, _ = torch.sort(torch.cat([z_free, z_surf], dim=-1), dim=-1)
    if t_rand is not None:
        z = perturb(z, t_rand)
    return z


def sample_pdf(bins, weights, u):
    """common.sample_pdf (common.py:49-85), det=False, with u = torch.rand passed in.
    Quirk kept: pdf = weights (UNnormalised, line 56 overwrites line 55). Returns samples, inds."""
    pdf = weights
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
    inds_g = torch.stack([below, above], -1)
    matched_shape = [inds_g.shape[0], inds_g.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(matched_shape), 2, inds_g)
    bins_g = torch.gather(bins.unsqueeze(1).expand(matched_shape), 2, inds_g)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    samples = bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])
    return samples, inds


# ----------------------------------------------------------------------------------------------
# field: hash grids + decoders  (src/networks/decoders.py:91-205)
# ----------------------------------------------------------------------------------------------
@dataclass
class Field:
    """scene_rep + Decoders state. variant 'A' = nn.Linear stacks (decoders.py:72-84),
    'B' = tcnn FullyFusedMLP restated in fp32 (decoders.py:49-70; 32->16 ReLU->16(padded) no bias)."""
    sdf_spec: grid_ref.GridSpec
    rgb_spec: grid_ref.GridSpec
    sdf_table: torch.Tensor
    rgb_table: torch.Tensor
    variant: str
    w: Dict[str, torch.Tensor]         # decoder weights (see init_decoder_weights)
    beta: torch.Tensor                 # Parameter[1] init 10 (decoders.py:86-89)
    bound: torch.Tensor                # (3,2)
    tape: Optional[list] = None        # when a list: every differentiable encode appends (grid, clamped points, features) with
                                       # features.retain_grad(), so a test can re-accumulate the table gradient in fp64

    def parameters(self):
        return [self.sdf_table, self.rgb_table, self.beta] + list(self.w.values())


def init_decoder_weights(variant: str, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Variant A: torch.nn.Linear default init, state_dict names of decoders.py:74-84.
    Variant B: flat 'params' of 768 floats per network: W1 (16,32) row-major then Wout (16,16)
    row-major (padded; first n_out rows used), xavier-uniform like tcnn."""
    g = torch.Generator().manual_seed(seed)

    def lin(o, i):
        k = 1.0 / (i ** 0.5)
        return ((torch.rand(o, i, generator=g) * 2 - 1) * k).to(dtype), ((torch.rand(o, generator=g) * 2 - 1) * k).to(dtype)

    w = {}
    if variant == "A":
        w["linears.0.weight"], w["linears.0.bias"] = lin(16, 32)
        w["linears.1.weight"], w["linears.1.bias"] = lin(16, 16)
        w["c_linears.0.weight"], w["c_linears.0.bias"] = lin(16, 32)
        w["c_linears.1.weight"], w["c_linears.1.bias"] = lin(16, 16)
        w["output_linear.weight"], w["output_linear.bias"] = lin(1, 16)
        w["c_output_linear.weight"], w["c_output_linear.bias"] = lin(3, 16)
    elif variant == "B":
        def xav(o, i):
            s = (6.0 / (i + o)) ** 0.5
            return ((torch.rand(o, i, generator=g) * 2 - 1) * s).to(dtype)
        w["sdf_decoder.params"] = torch.cat([xav(16, 32).reshape(-1), xav(16, 16).reshape(-1)])
        w["color_decoder.params"] = torch.cat([xav(16, 32).reshape(-1), xav(16, 16).reshape(-1)])
    else:
        raise ValueError(variant)
    return w


def mlp_B(h, params, n_out, out_act):
    """tcnn FullyFusedMLP(n_neurons=16, n_hidden_layers=1, ReLU, no bias) restated in fp32."""
    W1 = params[:512].reshape(16, 32)
    Wo = params[512:768].reshape(16, 16)[:n_out]
    h = torch.relu(h @ W1.t())
    return out_act(h @ Wo.t())


def raw_sdf(field: Field, p_nor):
    """Decoders.get_raw_sdf (decoders.py:107-130) incl. the clamp of sample_hash_grid_feature (:101)."""
    p = torch.clamp(p_nor, min=0, max=1)
    h = grid_ref.encode(field.sdf_spec, field.sdf_table, p)
    if field.tape is not None and h.requires_grad:
        h.retain_grad(); field.tape.append(("sdf", p.detach(), h))
    w = field.w
    if field.variant == "B":
        return mlp_B(h, w["sdf_decoder.params"], 1, torch.tanh).squeeze()
    h = torch.relu(F.linear(h, w["linears.0.weight"], w["linears.0.bias"]))
    h = torch.relu(F.linear(h, w["linears.1.weight"], w["linears.1.bias"]))
    return torch.tanh(F.linear(h, w["output_linear.weight"], w["output_linear.bias"])).squeeze()


def raw_rgb(field: Field, p_nor):
    """Decoders.get_raw_rgb (decoders.py:132-155)."""
    p = torch.clamp(p_nor, min=0, max=1)
    h = grid_ref.encode(field.rgb_spec, field.rgb_table, p)
    if field.tape is not None and h.requires_grad:
        h.retain_grad(); field.tape.append(("rgb", p.detach(), h))
    w = field.w
    if field.variant == "B":
        return mlp_B(h, w["color_decoder.params"], 3, torch.sigmoid)
    h = torch.relu(F.linear(h, w["c_linears.0.weight"], w["c_linears.0.bias"]))
    h = torch.relu(F.linear(h, w["c_linears.1.weight"], w["c_linears.1.bias"]))
    return torch.sigmoid(F.linear(h, w["c_output_linear.weight"], w["c_output_linear.bias"]))


def decoders_forward(field: Field, p):
    """Decoders.forward (decoders.py:182-205): raw = cat([rgb, sdf]) reshaped (...,4)."""
    shape = p.shape
    p_nor = p.reshape(-1, 3)
    sdf = raw_sdf(field, p_nor)
    rgb = raw_rgb(field, p_nor)
    raw = torch.cat([rgb, sdf.reshape(-1, 1)], dim=-1)
    return raw.reshape(*shape[:-1], -1)


# ----------------------------------------------------------------------------------------------
# compositing  (src/utils/Renderer.py:132-158)
# ----------------------------------------------------------------------------------------------
def sdf2alpha(sdf, beta):
    """Renderer.sdf2alpha (Renderer.py:154-158)."""
    return 1. - torch.exp(-beta * torch.sigmoid(-sdf * beta))


def composite(raw, z_vals, beta):
    """Renderer.py:140-152. Returns the 7-tuple of render_batch_ray."""
    alpha = sdf2alpha(raw[..., 3], beta)
    ones = torch.ones((alpha.shape[0], 1), device=alpha.device, dtype=alpha.dtype)
    weights = alpha * torch.cumprod(torch.cat([ones, (1. - alpha + 1e-10)], -1), -1)[:, :-1]
    rgb = torch.sum(weights[..., None] * raw[..., :3], -2)
    depth = torch.sum(weights * z_vals, -1)
    term = torch.sum(weights, -1)
    pixel_unc = torch.square(1 - torch.sum(weights, -1))
    depth_unc = torch.sqrt(torch.sum(weights * (depth[..., None] - z_vals) ** 2, -1))
    return term, pixel_unc, depth, rgb, raw[..., 3], z_vals, depth_unc


def zvals_no_depth(field: Field, rays_o, rays_d, n_stratified, n_importance, t_rand_uni, u_pdf):
    """Renderer.py:103-130 (rays with gt_depth == 0; no_grad). Quirk kept: coordinates are
    normalised to [-1,1] (common.normalize_3d_coordinate) and then clamped to [0,1].
    t_rand_uni (R0,n_strat) or None; u_pdf (R0,n_importance). Returns z (R0,S) and the
    searchsorted indices of sample_pdf."""
    with torch.no_grad():
        dev = rays_o.device
        bound = field.bound
        t_uni = torch.linspace(0., 1., steps=n_stratified, device=dev)
        far_bb = bbox_exit(rays_o, rays_d, bound).unsqueeze(-1)
        far_bb = far_bb + 0.01
        z_uni = 0.0 * (1. - t_uni) + far_bb * t_uni
        if t_rand_uni is not None:
            z_uni = perturb(z_uni, t_rand_uni)
        pts = rays_o.detach().unsqueeze(1) + rays_d.detach().unsqueeze(1) * z_uni.unsqueeze(-1)
        p = pts.reshape(-1, 3).clone()
        for a in range(3):
            p[:, a] = ((p[:, a] - bound[a, 0]) / (bound[a, 1] - bound[a, 0])) * 2 - 1.0
        sdf = raw_sdf(field, p).reshape(*pts.shape[0:2])
        alpha = sdf2alpha(sdf, field.beta)
        ones = torch.ones((alpha.shape[0], 1), device=dev, dtype=alpha.dtype)
        weights = alpha * torch.cumprod(torch.cat([ones, (1. - alpha + 1e-10)], -1), -1)[:, :-1]
        z_mid = .5 * (z_uni[..., 1:] + z_uni[..., :-1])
        z_samples, inds = sample_pdf(z_mid, weights[..., 1:-1], u_pdf)
        z, _ = torch.sort(torch.cat([z_uni, z_samples], -1), -1)
    return z, inds


def render_batch_ray(field: Field, rays_d, rays_o, gt_depth, n_stratified, n_importance, truncation,
                     t_rand=None, t_rand_uni=None, u_pdf=None, parts: Optional[dict] = None, nodepth_field: Optional[Field] = None,
                     z_nodepth_override=None):
    """Renderer.render_batch_ray (Renderer.py:59-152) with RNG draws passed in:
    t_rand (R_valid,S): perturbation of depth-guided rays; t_rand_uni (R0,n_strat), u_pdf (R0,n_imp):
    draws of the no-depth branch, in the order the reference consumes them.
    parts (optional) receives 'pdf_inds' (R0,n_imp): the torch.searchsorted indices of sample_pdf (common.py:70).
    nodepth_field (optional): field used for the no_grad z-sampling of depth-less rays -- an fp64 gradient check passes
    the fp32 field here so that both runs integrate along identical sample positions.
    z_nodepth_override (optional, (R0,S)): sample positions to integrate the depth-less rays along INSTEAD of the ones computed
    here (which parts['z_nodepth_own'] then records).  A stage-wise parity check uses it: the inverse-cdf resampling amplifies
    ulp-level SDF differences into ~1e-5 relative shifts of z, so the checked implementation's z is first held to its own
    bar against z_nodepth_own and then fed in, and everything downstream is compared along identical positions."""
    n_rays = rays_o.shape[0]
    S = n_stratified + n_importance
    z_vals = torch.empty([n_rays, S], device=rays_o.device, dtype=torch.float32)
    gt_depth = gt_depth.reshape(-1, 1)
    gt_mask = (gt_depth > 0).squeeze(-1)
    z_vals[gt_mask] = zvals_with_depth(gt_depth[gt_mask], n_stratified, n_importance, truncation, t_rand)
    if not gt_mask.all():
        z0, inds = zvals_no_depth(nodepth_field if nodepth_field is not None else field, rays_o[~gt_mask], rays_d[~gt_mask],
                                  n_stratified, n_importance, t_rand_uni, u_pdf)
        if parts is not None:
            parts["pdf_inds"] = inds
            parts["z_nodepth_own"] = z0.to(z_vals.dtype)
        z_vals[~gt_mask] = z0.to(z_vals.dtype) if z_nodepth_override is None else z_nodepth_override.to(z_vals.dtype)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    bound = field.bound
    pts = (pts - bound[:, 0]) / (bound[:, 1] - bound[:, 0])
    if field.sdf_table.dtype == torch.float64:
        pts = pts.double(); z_vals = z_vals.double()
    raw = decoders_forward(field, pts)
    return composite(raw, z_vals, field.beta)


def render_img(field: Field, H, W, fx, fy, cx, cy, c2w, gt_depth, n_stratified, n_importance, truncation,
               ray_batch_size, draw_rand):
    """Renderer.render_img (Renderer.py:160-223): get_rays over the whole frame, render_batch_ray per chunk of
    ray_batch_size rays (the last one ragged), outputs concatenated; everything but colour is returned as float64
    (Renderer.py:205-209).  draw_rand(shape) supplies torch.rand draws in the reference's consumption order: per chunk
    (n_valid,S), then -- only if the chunk has depth-less rays -- (n0,n_stratified) and (n0,n_importance).
    Returns (depth, color, termination_prob, pixel_unc, depth_unc) shaped (H,W[,3])."""
    S = n_stratified + n_importance
    with torch.no_grad():
        rays_o, rays_d = full_image_rays(H, W, fx, fy, cx, cy, c2w)
        rays_o = rays_o.reshape(-1, 3); rays_d = rays_d.reshape(-1, 3)
        gt_depth = gt_depth.reshape(-1)
        outs = [[] for _ in range(5)]
        for i in range(0, rays_d.shape[0], ray_batch_size):
            gt = gt_depth[i:i + ray_batch_size]
            n_valid = int((gt > 0).sum()); n0 = gt.numel() - n_valid
            t_rand = draw_rand((n_valid, S))
            t_uni = draw_rand((n0, n_stratified)) if n0 > 0 else None
            u_pdf = draw_rand((n0, n_importance)) if n0 > 0 else None
            term, punc, depth, color, _, _, dunc = render_batch_ray(field, rays_d[i:i + ray_batch_size], rays_o[i:i + ray_batch_size], gt,
                                                                    n_stratified, n_importance, truncation, t_rand, t_uni, u_pdf)
            for lst, v in zip(outs, (depth.double(), color, term.double(), punc.double(), dunc.double())):
                lst.append(v)
        depth, color, term, punc, dunc = [torch.cat(o, dim=0) for o in outs]
        return depth.reshape(H, W), color.reshape(H, W, 3), term.reshape(H, W), punc.reshape(H, W), dunc.reshape(H, W)


# ----------------------------------------------------------------------------------------------
# losses  (src/Mapper.py:141-175,412-440; src/Tracker.py:113-147,208-238)
# ----------------------------------------------------------------------------------------------
@dataclass
class LossWeights:
    w_sdf_fs: float
    w_sdf_center: float
    w_sdf_tail: float
    w_depth: float
    w_color: float


MAP_WEIGHTS = LossWeights(5, 200, 10, 0.1, 5)        # configs/UNISLAM.yaml:67-71
TRACK_WEIGHTS = LossWeights(10, 200, 50, 1, 5)       # configs/UNISLAM.yaml:42-46


def sdf_losses(sdf, z_vals, gt_depth, truncation, lw: LossWeights, parts: Optional[dict] = None):
    """Mapper.sdf_losses == Tracker.sdf_losses (Mapper.py:141-175). Thresholds are evaluated in
    fp32 exactly as the reference does (the masks depend only on z and gt)."""
    z32 = z_vals.to(torch.float32); g32 = gt_depth.to(torch.float32)
    front = z32 < (g32[:, None] - truncation)
    back = z32 > (g32[:, None] + truncation)
    center = (z32 > (g32[:, None] - 0.4 * truncation)) & (z32 < (g32[:, None] + 0.4 * truncation))
    tail = (~front) & (~back) & (~center)
    gt_e = gt_depth[:, None].expand(z_vals.shape)
    fs = torch.mean(torch.square(sdf[front] - torch.ones_like(sdf[front])))
    ce = torch.mean(torch.square((z_vals + sdf * truncation)[center] - gt_e[center]))
    ta = torch.mean(torch.square((z_vals + sdf * truncation)[tail] - gt_e[tail]))
    if parts is not None:
        parts.update(fs=fs, center=ce, tail=ta, n_front=int(front.sum()), n_center=int(center.sum()),
                     n_tail=int(tail.sum()))
    return lw.w_sdf_fs * fs + lw.w_sdf_center * ce + lw.w_sdf_tail * ta


def _no_mask_loss(ret, gt_depth, gt_color, truncation, lw, parts):
    """m_mask_mode / t_mask_mode == "no_mask" (Mapper.py:432-440, Tracker.py:230-238): every term over every ray.
    (The tracker writes (w*x).mean() where the mapper writes w*x.mean(): the same number.)"""
    term, pixel_unc, depth, color, sdf, z_vals, _ = ret
    gtd = gt_depth.to(depth.dtype)
    loss = sdf_losses(sdf, z_vals, gtd, truncation, lw, parts)
    col = torch.square(gt_color.to(color.dtype) - color).mean()
    dep = torch.square(gtd - depth).mean()
    if parts is not None:
        parts.update(color=col, depth=dep, n_mask=int(gtd.numel()), depth_mask=torch.ones_like(gtd, dtype=torch.bool))
    return loss + lw.w_color * col + lw.w_depth * dep


def mapping_loss(ret, gt_depth, gt_color, truncation, lw: LossWeights = MAP_WEIGHTS, parts=None, mask_mode="original"):
    """Mapper.py:412-440."""
    if mask_mode == "no_mask":
        return _no_mask_loss(ret, gt_depth, gt_color, truncation, lw, parts)
    term, pixel_unc, depth, color, sdf, z_vals, _ = ret
    alpha_mask = (1 - pixel_unc.detach()).to(torch.float32) > 0.99
    depth_mask = (gt_depth > 0) & alpha_mask
    gtd = gt_depth.to(depth.dtype)
    loss = sdf_losses(sdf[depth_mask], z_vals[depth_mask], gtd[depth_mask], truncation, lw, parts)
    col = torch.square(gt_color.to(color.dtype) - color).mean()
    dep = torch.square(gtd[depth_mask] - depth[depth_mask]).mean()
    if parts is not None:
        parts.update(color=col, depth=dep, n_mask=int(depth_mask.sum()), depth_mask=depth_mask)
    return loss + lw.w_color * col + lw.w_depth * dep


def tracking_loss(ret, gt_depth, gt_color, truncation, lw: LossWeights = TRACK_WEIGHTS, parts=None, mask_mode="original"):
    """Tracker.py:208-238. 'original': 10x-median depth-error mask, colour over masked rays."""
    if mask_mode == "no_mask":
        return _no_mask_loss(ret, gt_depth, gt_color, truncation, lw, parts)
    term, pixel_unc, depth, color, sdf, z_vals, _ = ret
    alpha_mask = (1 - pixel_unc.detach()).to(torch.float32) > 0.99
    gtd = gt_depth.to(depth.dtype)
    depth_error = (gtd - depth.detach()).abs()
    error_median = depth_error.median()
    depth_mask = (depth_error < 10 * error_median) & alpha_mask
    loss = sdf_losses(sdf[depth_mask], z_vals[depth_mask], gtd[depth_mask], truncation, lw, parts)
    col = torch.square(gt_color.to(color.dtype) - color)[depth_mask].mean()
    dep = torch.square(gtd[depth_mask] - depth[depth_mask]).mean()
    if parts is not None:
        parts.update(color=col, depth=dep, n_mask=int(depth_mask.sum()), depth_mask=depth_mask,
                     median=error_median)
    return loss + lw.w_color * col + lw.w_depth * dep


# ----------------------------------------------------------------------------------------------
# whole iterations (sample -> prefilter -> render -> loss), RNG passed in
# ----------------------------------------------------------------------------------------------
def mapping_iteration(field: Field, batches, truncation, n_stratified, n_importance, draw_rand,
                      lw: LossWeights = MAP_WEIGHTS, parts=None, mask_mode="original", nodepth_field: Optional[Field] = None,
                      z_nodepth_override=None):
    """Mapper.optimize_mapping body, Mapper.py:379-430. batches = [(c2ws, depths, colors, rays_d_cam,
    indices), ...]: the main get_samples_all call (:379) and, when >20 keyframes, the 200 px x last-10
    frames call (:385-393), concatenated in that order. draw_rand(shape) supplies the torch.rand
    draws of render_batch_ray in reference order."""
    outs = [sample_mapping_rays(*b) for b in batches]
    rays_o, rays_d, gt_depth, gt_color = [torch.cat([o[i] for o in outs], dim=0) for i in range(4)]
    with torch.no_grad():
        inside = bbox_exit(rays_o, rays_d, field.bound) >= gt_depth
    rays_d, rays_o, gt_depth, gt_color = rays_d[inside], rays_o[inside], gt_depth[inside], gt_color[inside]
    S = n_stratified + n_importance
    n_valid = int((gt_depth > 0).sum()); n0 = gt_depth.shape[0] - n_valid
    t_rand = draw_rand((n_valid, S))
    t_uni = draw_rand((n0, n_stratified)) if n0 > 0 else None
    u_pdf = draw_rand((n0, n_importance)) if n0 > 0 else None
    ret = render_batch_ray(field, rays_d, rays_o, gt_depth, n_stratified, n_importance, truncation,
                           t_rand, t_uni, u_pdf, parts, nodepth_field, z_nodepth_override)
    if parts is not None:
        parts.update(inside=inside, ret=ret, gt_depth=gt_depth, gt_color=gt_color, rays_o=rays_o, rays_d=rays_d)
    return mapping_loss(ret, gt_depth, gt_color, truncation, lw, parts, mask_mode)


def tracking_iteration(field: Field, cam_pose, depth_img, color_img, H, W, fx, fy, cx, cy, edge_h, edge_w,
                       indices, truncation, n_stratified, n_importance, draw_rand,
                       lw: LossWeights = TRACK_WEIGHTS, parts=None, mask_mode="original"):
    """Tracker.optimize_tracking, Tracker.py:170-238. cam_pose (1,7) requires grad."""
    c2w = cam_pose_to_matrix(cam_pose)
    rays_o, rays_d, gt_all, gt_color = sample_tracking_rays(edge_h, H - edge_h, edge_w, W - edge_w, indices.numel(),
                                                            fx, fy, cx, cy, c2w, depth_img, color_img, indices)
    with torch.no_grad():
        inside = (bbox_exit(rays_o, rays_d, field.bound) >= gt_all) & (gt_all > 0)
    rays_d, rays_o, gt_depth, gt_color = rays_d[inside], rays_o[inside], gt_all[inside], gt_color[inside]
    S = n_stratified + n_importance
    t_rand = draw_rand((gt_depth.shape[0], S))
    ret = render_batch_ray(field, rays_d, rays_o, gt_depth, n_stratified, n_importance, truncation, t_rand)
    if parts is not None:
        parts.update(inside=inside, ret=ret, gt_depth=gt_depth, gt_color=gt_color, rays_o=rays_o, rays_d=rays_d)
    return tracking_loss(ret, gt_depth, gt_color, truncation, lw, parts, mask_mode), ret[1]


# ----------------------------------------------------------------------------------------------
# keyframe co-visibility  (src/Mapper.py:177-236)
# ----------------------------------------------------------------------------------------------
def keyframe_covisibility(rays_o, rays_d, gt_depth, keyframe_c2ws, H, W, fx, fy, cx, cy, num_samples=8, edge=20):
    """percent_inside of Mapper.keyframe_selection_LC (Mapper.py:199-236): points sampled in [0.8 d, d + 0.5] along the rays
    with sensor depth, projected into every keyframe (the caller drops the last two, Mapper.py:214)."""
    dev = rays_o.device
    gt_depth = gt_depth.reshape(-1, 1)
    nz = gt_depth[:, 0] > 0
    rays_o, rays_d, gt_depth = rays_o[nz], rays_d[nz], gt_depth[nz].repeat(1, num_samples)
    t_vals = torch.linspace(0., 1., steps=num_samples).to(dev)
    z_vals = gt_depth * 0.8 * (1. - t_vals) + (gt_depth + 0.5) * t_vals
    pts = (rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]).reshape(1, -1, 3)
    w2cs = torch.inverse(keyframe_c2ws)
    ones = torch.ones_like(pts[..., 0]).reshape(1, -1, 1)
    homo = torch.cat([pts, ones], dim=-1).reshape(1, -1, 4, 1).expand(w2cs.shape[0], -1, -1, -1)
    cam = (w2cs.unsqueeze(1).expand(-1, homo.shape[1], -1, -1) @ homo)[:, :, :3]
    K = torch.tensor([[fx, .0, cx], [.0, fy, cy], [.0, .0, 1.0]], device=dev).reshape(3, 3)
    cam[:, :, 0] *= -1
    uv = K @ cam
    z = uv[:, :, -1:] + 1e-5
    uv = uv[:, :, :2] / z
    mask = (uv[:, :, 0] < W - edge) * (uv[:, :, 0] > edge) * (uv[:, :, 1] < H - edge) * (uv[:, :, 1] > edge)
    mask = (mask & (z[:, :, 0] < 0)).squeeze(-1)
    return mask.sum(dim=1) / uv.shape[1]


# ----------------------------------------------------------------------------------------------
# dense SDF query for meshing  (src/utils/Mesher.py:134-195)
# ----------------------------------------------------------------------------------------------
def mesh_grid_axes(mc_bound, resolution=0.01, padding=0.05):
    """Mesher.get_grid_uniform (Mesher.py:168-195): per-axis np.linspace -> fp32."""
    import numpy as np
    axes = []
    for a in range(3):
        lo, hi = float(mc_bound[a][0]), float(mc_bound[a][1])
        n = int(torch.tensor((hi - lo + 2 * padding) / resolution).round().int().item())
        axes.append(torch.from_numpy(np.linspace(lo - padding, hi + padding, n)).float())
    return axes


def mesh_grid_points(axes):
    gx, gy, gz = torch.meshgrid(axes[0], axes[1], axes[2], indexing="xy")
    return torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], dim=1)


def eval_points_sdf(field: Field, p):
    """Mesher.eval_points (Mesher.py:134-166), SDF channel only: strict in-bound mask, -1 outside."""
    b = field.bound.to(p)
    mask = ((p[:, 0] < b[0][1]) & (p[:, 0] > b[0][0]) & (p[:, 1] < b[1][1]) & (p[:, 1] > b[1][0])
            & (p[:, 2] < b[2][1]) & (p[:, 2] > b[2][0]))
    pn = (p - b[:, 0]) / (b[:, 1] - b[:, 0])
    sdf = raw_sdf(field, pn).clone()
    sdf[~mask] = -1
    return sdf


def load_bound(bound_yaml, scale=1.0, bound_dividable=0.24):
    """UNISLAM.load_bound (UNISLAM.py:205-222)."""
    import numpy as np
    bound = torch.from_numpy(np.array(bound_yaml) * scale).float()
    bound[:, 1] = (((bound[:, 1] - bound[:, 0]) / bound_dividable).int() + 1) * bound_dividable + bound[:, 0]
    return bound


# ---- eval_rendering's per-frame metrics (src/tools/eval_recon.py:276-286) -----------------------------------------------------
def render_metrics(gt_color, gt_depth, color, depth):
    """(mse, psnr, depth_l1) of one rendered frame over the pixels with sensor depth: mse_loss(gt_color[m], color[m]) with the
    dataset's float64 colour promoting the difference to float64, psnr = -10 log10(mse), depth_l1 = mean |gt_depth[m] - depth[m]|
    (render_img returns depth as float64)."""
    m = gt_depth > 0
    mse = ((gt_color[m].double() - color[m].double()) ** 2).mean()
    l1 = (gt_depth[m].double() - depth[m].double()).abs().mean()
    return mse, -10.0 * torch.log10(mse), l1
