"""Host-side mirror of the reference's operator interface for the hot path.

* ``Encoding`` / ``Network``  -- drop-ins for ``tinycudann.Encoding`` / ``tinycudann.Network`` as
  constructed at src/UNISLAM.py:242-253 and src/networks/decoders.py:50-70 (seams B1, B2).
* ``Decoders``                -- same constructor, attributes, state_dict keys and methods as
  src/networks/decoders.py:24-205 (seam B3), evaluated by the fused field kernels.
* ``Renderer``                -- ``render_batch_ray`` / ``render_img`` with the signature and RNG
  consumption order of src/utils/Renderer.py:59-223 (seam B4).

All arithmetic runs in lib/libunislam_b200.so; nothing here has a CPU path.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops


# ------------------------------------------------------------------------------------------------
class Encoding(nn.Module):
    """tinycudann.Encoding(n_input_dims=3, encoding_config={otype: HashGrid, ...}, dtype=torch.float).

    ``params`` is one flat fp32 leaf Parameter [sum_l T_l * 2] (level-major, entry-major, feature-minor),
    initialised U(-1e-4, 1e-4) like tcnn; picklable across mp.spawn and deep-copyable (the native grid
    descriptor is rebuilt from the config)."""

    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__()
        if n_input_dims != 3:
            raise RuntimeError("unislam_b200.Encoding: only n_input_dims=3 is supported")
        cfg = dict(encoding_config)
        if cfg.get("otype", "HashGrid") not in ("HashGrid", "Grid"):
            raise RuntimeError(f"unislam_b200.Encoding: unsupported otype {cfg.get('otype')}")
        if int(cfg.get("n_features_per_level", 2)) != 2 or int(cfg.get("n_levels", 16)) > L.MAX_LEVELS:
            raise RuntimeError("unislam_b200.Encoding: needs n_features_per_level=2 and n_levels<=16")
        if dtype not in (None, torch.float, torch.float32):
            raise RuntimeError("unislam_b200.Encoding: only dtype=torch.float (fp32 tables) is supported")
        self.n_input_dims = 3
        self.encoding_config = cfg
        self.seed = seed
        self._build()
        g = torch.Generator().manual_seed(seed)
        init = (torch.rand(self.grid.total_entries * 2, generator=g) * 2 - 1) * 1e-4
        self.params = nn.Parameter(init)

    def _build(self):
        c = self.encoding_config
        self.grid = L.build_grid(int(c.get("n_levels", 16)), int(c.get("log2_hashmap_size", 19)),
                                 int(c.get("base_resolution", 16)), float(c.get("per_level_scale", 2.0)))
        self.n_output_dims = self.grid.n_levels * 2

    def __getstate__(self):
        st = dict(self.__dict__)
        st.pop("grid", None)
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._build()

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        nn.Module.__init__(new)
        new.n_input_dims, new.encoding_config, new.seed = 3, dict(self.encoding_config), self.seed
        new._build()
        new.params = nn.Parameter(self.params.detach().clone(), requires_grad=self.params.requires_grad)
        memo[id(self)] = new
        return new

    def level_table(self):
        return [(lv.scale, lv.res, lv.size, lv.offset, bool(lv.hashed)) for lv in list(self.grid.levels)[: self.grid.n_levels]]

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("unislam_b200.Encoding: input must be a CUDA tensor (no CPU fallback)")
        shape = x.shape
        y = ops.grid_encode(x.reshape(-1, 3), self.params, self.grid)
        return y.reshape(*shape[:-1], self.n_output_dims)

    def extra_repr(self):
        return f"n_input_dims=3, n_output_dims={self.n_output_dims}, config={self.encoding_config}"


class _MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layout: ops.DecoderLayout, h, *tensors):
        h = L.f32c(h)
        n = h.shape[0]
        out = torch.empty((n, layout.n_out), device=h.device, dtype=torch.float32)
        m = layout.pack(tensors)
        from ctypes import byref
        L.call("usl_mlp_fwd", byref(m), L.ptr(h), n, L.ptr(out), L.stream())
        ctx.layout = layout
        ctx.save_for_backward(h, out, *tensors)
        return out

    @staticmethod
    def backward(ctx, dout):
        from ctypes import byref
        h, out, *tensors = ctx.saved_tensors
        layout = ctx.layout
        dout = L.f32c(dout)
        n = h.shape[0]
        need_w = any(ctx.needs_input_grad[2:])
        grads = [torch.zeros_like(t) for t in tensors] if need_w else [None] * len(tensors)
        dh = torch.empty_like(h) if ctx.needs_input_grad[1] else None
        m = layout.pack(tensors)
        gm = layout.pack(grads) if need_w else None
        L.call("usl_mlp_bwd", byref(m), byref(gm) if gm is not None else None, L.ptr(h), L.ptr(out), L.ptr(dout), n,
               L.ptr(dh), L.stream())
        return (None, dh, *grads)


class Network(nn.Module):
    """tinycudann.Network(n_input_dims=32, n_output_dims, {FullyFusedMLP, ReLU, Tanh|Sigmoid, n_neurons 16,
    n_hidden_layers 1}) restated in fp32 (the real one computes in fp16; parity target is the fp32
    restatement). ``params``: 768 fp32 = W1 (16,32) then Wout (16,16) row-major."""

    _ACTS = {"Tanh": L.ACT_TANH, "Sigmoid": L.ACT_SIGMOID, "None": L.ACT_NONE}

    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        c = dict(network_config)
        if (n_input_dims != ops.C_DIM or c.get("n_neurons") != ops.HIDDEN or c.get("n_hidden_layers") != 1
                or c.get("activation") != "ReLU" or n_output_dims not in (1, 2, 3)
                or c.get("output_activation", "None") not in self._ACTS):
            raise RuntimeError("unislam_b200.Network: only 32 -> 16 (ReLU) -> {1..3} FullyFusedMLP decoders are supported")
        self.n_input_dims, self.n_output_dims, self.network_config, self.seed = n_input_dims, n_output_dims, c, seed
        g = torch.Generator().manual_seed(seed)
        xav = lambda o, i: (torch.rand(o, i, generator=g) * 2 - 1) * math.sqrt(6.0 / (i + o))
        self.params = nn.Parameter(torch.cat([xav(16, 32).reshape(-1), xav(16, 16).reshape(-1)]))

    @property
    def layout(self):
        return ops.DecoderLayout("B", self.n_output_dims, self._ACTS[self.network_config.get("output_activation", "None")])

    def forward(self, h):
        if not h.is_cuda:
            raise RuntimeError("unislam_b200.Network: input must be a CUDA tensor (no CPU fallback)")
        return _MlpFn.apply(self.layout, h.reshape(-1, ops.C_DIM), self.params)


# ------------------------------------------------------------------------------------------------
class Decoders(nn.Module):
    """Drop-in for src/networks/decoders.py:Decoders with the field query fused into one launch per
    direction. Same constructor signature, ``beta`` / ``bound`` attributes and state_dict keys
    (``linears.*``, ``c_linears.*``, ``output_linear.*``, ``c_output_linear.*`` | ``sdf_decoder.params``,
    ``color_decoder.params``; ``beta``)."""

    def __init__(self, cfg, c_dim=32, hidden_size=16, truncation=0.08, n_blocks=2, learnable_beta=True):
        super().__init__()
        if c_dim != ops.C_DIM or hidden_size != ops.HIDDEN or n_blocks != 2 or cfg.get("grid_mode", "hash_grid") != "hash_grid":
            raise RuntimeError("unislam_b200.Decoders: only c_dim=32, hidden_size=16, n_blocks=2, grid_mode=hash_grid")
        self.c_dim, self.cfg, self.truncation, self.n_blocks = c_dim, cfg, truncation, n_blocks
        self.tcnn_network = bool(cfg["grid"]["tcnn_network"])
        if self.tcnn_network:
            net = {"otype": "FullyFusedMLP", "activation": "ReLU", "n_neurons": hidden_size, "n_hidden_layers": n_blocks - 1}
            self.sdf_decoder = Network(c_dim, 1, dict(net, output_activation="Tanh"))
            self.color_decoder = Network(c_dim, 3, dict(net, output_activation="Sigmoid"))
        else:
            self.linears = nn.ModuleList([nn.Linear(c_dim, hidden_size)] + [nn.Linear(hidden_size, hidden_size) for _ in range(n_blocks - 1)])
            self.c_linears = nn.ModuleList([nn.Linear(c_dim, hidden_size)] + [nn.Linear(hidden_size, hidden_size) for _ in range(n_blocks - 1)])
            self.output_linear = nn.Linear(hidden_size, 1)
            self.c_output_linear = nn.Linear(hidden_size, 3)
        if learnable_beta:
            self.beta = nn.Parameter(10 * torch.ones(1))
        else:
            self.beta = 10
        self.bound = None
        self._beta_const = None

    # -- plumbing ----------------------------------------------------------------------------
    @property
    def variant(self):
        return "B" if self.tcnn_network else "A"

    def decoder_tensors(self):
        if self.tcnn_network:
            return [self.sdf_decoder.params, self.color_decoder.params]
        return [self.linears[0].weight, self.linears[0].bias, self.linears[1].weight, self.linears[1].bias,
                self.output_linear.weight, self.output_linear.bias,
                self.c_linears[0].weight, self.c_linears[0].bias, self.c_linears[1].weight, self.c_linears[1].bias,
                self.c_output_linear.weight, self.c_output_linear.bias]

    def beta_tensor(self, device):
        if isinstance(self.beta, torch.Tensor):
            return self.beta
        if self._beta_const is None or self._beta_const.device != torch.device(device):
            self._beta_const = torch.full((1,), float(self.beta), device=device)
        return self._beta_const

    def field_meta(self, scene_rep, bound=None) -> ops.FieldMeta:
        grids, c_grids = scene_rep
        b = bound if bound is not None else self.bound
        if b is None:
            # points handed to Decoders.forward are already normalised: identity bound
            b = torch.tensor([[0., 1.]] * 3)
        return ops.FieldMeta(grids[0].grid, c_grids[0].grid, self.variant, L.make_bound(b))

    # -- reference API -----------------------------------------------------------------------
    def get_raw_sdf(self, p_nor, scene_rep):
        """decoders.py:107-130 (forward only: its sole caller runs under no_grad, Renderer.py:105-121)."""
        grids, c_grids = scene_rep
        meta = self.field_meta(scene_rep)
        out = ops.field_sdf_points(meta, p_nor.reshape(-1, 3), grids[0].params, c_grids[0].params, self.decoder_tensors())
        return out.squeeze()

    def forward(self, p, scene_rep):
        """decoders.py:182-205: raw (...,4) = (r,g,b,sdf) of already-normalised points."""
        grids, c_grids = scene_rep
        shape = p.shape
        meta = self.field_meta(scene_rep)
        raw = ops.field_points(meta, p.reshape(-1, 3), grids[0].params, c_grids[0].params, self.decoder_tensors())
        return raw.reshape(*shape[:-1], 4)


# ------------------------------------------------------------------------------------------------
class Renderer(object):
    """Drop-in for src/utils/Renderer.py:Renderer. ``render_batch_ray`` returns the same 7-tuple, is
    differentiable wrt rays, tables, decoder weights and beta, and consumes torch.rand in the reference's
    order: (n_valid,S) for depth-guided rays, then (n0,n_stratified) and (n0,n_importance) for no-depth rays."""

    def __init__(self, cfg, unislam, ray_batch_size=10000):
        self.ray_batch_size = ray_batch_size
        self.cfg = cfg
        self.perturb = cfg["rendering"]["perturb"]
        self.n_stratified = cfg["rendering"]["n_stratified"]
        self.n_importance = cfg["rendering"]["n_importance"]
        self.scale = cfg["scale"]
        self.bound = unislam.bound.to(unislam.device, non_blocking=True)
        self.H, self.W, self.fx, self.fy, self.cx, self.cy = unislam.H, unislam.W, unislam.fx, unislam.fy, unislam.cx, unislam.cy
        self._bound_c = L.make_bound(unislam.bound)
        self._zs = {}

    def _zsampler(self, truncation, device):
        key = (float(truncation), str(device))
        if key not in self._zs:
            self._zs[key] = ops.ZSampler(self.n_stratified, self.n_importance, float(truncation), device)
        return self._zs[key]

    def sdf2alpha(self, sdf, beta=10):
        return 1. - torch.exp(-beta * torch.sigmoid(-sdf * beta))

    def render_batch_ray(self, scene_rep, decoders, rays_d, rays_o, device, truncation, gt_depth=None):
        grids, c_grids = scene_rep
        R = rays_o.shape[0]
        zs = self._zsampler(truncation, rays_o.device)
        S = zs.S
        gt = L.f32c(gt_depth.reshape(-1))
        meta = ops.FieldMeta(grids[0].grid, c_grids[0].grid, decoders.variant, self._bound_c)
        dec = decoders.decoder_tensors()
        beta = decoders.beta_tensor(rays_o.device)
        z_vals = torch.empty((R, S), device=rays_o.device, dtype=torch.float32)
        gt_mask = gt > 0
        n_valid = int(gt_mask.sum())                     # same host sync as gt_depth[gt_mask] in the reference (Renderer.py:84)
        n0 = R - n_valid
        row_map = None
        if n0 > 0:
            # row of each ray inside the compacted (n_valid,.) / (n0,.) draws
            cv = torch.cumsum(gt_mask.to(torch.int32), 0) - 1
            c0 = torch.cumsum((~gt_mask).to(torch.int32), 0) - 1
            row_map = torch.where(gt_mask, cv, c0).to(torch.int32).contiguous()
        t_rand = torch.rand((n_valid, S), device=rays_o.device) if self.perturb else None
        zs.depth_guided(gt, z_vals, t_rand=t_rand, row_map=row_map)
        if n0 > 0:
            with torch.no_grad():
                t_uni = torch.rand((n0, self.n_stratified), device=rays_o.device) if self.perturb else None
                u_pdf = torch.rand((n0, self.n_importance), device=rays_o.device)
                f = meta.pack(grids[0].params, c_grids[0].params, dec)
                zs.no_depth(f, beta.detach(), L.f32c(rays_o.detach()), L.f32c(rays_d.detach()), gt, z_vals, u_pdf,
                            t_rand_uni=t_uni, row_map=row_map)
        term, punc, depth, rgb, sdf, dunc = ops.render_rays(meta, rays_o, rays_d, z_vals, beta, grids[0].params,
                                                            c_grids[0].params, dec)
        return term, punc, depth, rgb, sdf, z_vals, dunc

    def render_img(self, scene_rep, decoders, c2w, truncation, device, gt_depth=None):
        """Renderer.py:160-223: full frame in ray_batch_size chunks under no_grad; float64 outputs except colour."""
        with torch.no_grad():
            H, W = self.H, self.W
            if isinstance(c2w, np.ndarray):
                c2w = torch.from_numpy(c2w)
            rays_o, rays_d = ops.image_rays(c2w.to(device).float(), H, W, self.fx, self.fy, self.cx, self.cy)
            rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
            gt_depth = gt_depth.reshape(-1)
            outs = [[] for _ in range(5)]
            for i in range(0, rays_d.shape[0], self.ray_batch_size):
                ret = self.render_batch_ray(scene_rep, decoders, rays_d[i:i + self.ray_batch_size], rays_o[i:i + self.ray_batch_size],
                                            device, truncation, gt_depth=gt_depth[i:i + self.ray_batch_size])
                term, punc, depth, color, _, _, dunc = ret
                for lst, v in zip(outs, (term.double(), punc.double(), dunc.double(), depth.double(), color)):
                    lst.append(v)
            term, punc, dunc, depth, color = [torch.cat(o, dim=0) for o in outs]
            return depth.reshape(H, W), color.reshape(H, W, 3), term.reshape(H, W), punc.reshape(H, W), dunc.reshape(H, W)
