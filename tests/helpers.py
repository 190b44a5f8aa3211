"""Shared test helpers: golden loading, oracle Field construction (tests only)."""
import importlib
import os

import numpy as np
import torch

from oracle import grid_ref, path_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

DEC_SHAPES = {
    "A": {"c_linears.0.bias": (16,), "c_linears.0.weight": (16, 32), "c_linears.1.bias": (16,), "c_linears.1.weight": (16, 16),
          "c_output_linear.bias": (3,), "c_output_linear.weight": (3, 16), "linears.0.bias": (16,), "linears.0.weight": (16, 32),
          "linears.1.bias": (16,), "linears.1.weight": (16, 16), "output_linear.bias": (1,), "output_linear.weight": (1, 16)},
    "B": {"color_decoder.params": (768,), "sdf_decoder.params": (768,)},
}


def pkg():
    return importlib.import_module("uni-slam_b200")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def golden_decoder_weights(variant, seed_salt, dtype=torch.float32):
    """Same integer-hash fill as oracle/gen_golden.py:_build_world (sorted named_parameters order)."""
    names = sorted(list(DEC_SHAPES[variant].keys()) + ["beta"])
    w = {}
    for k, name in enumerate(names):
        if name == "beta":
            continue
        shp = DEC_SHAPES[variant][name]
        n = int(np.prod(shp))
        fan_in = shp[-1] if len(shp) > 1 else 16
        sc = 1.0 / np.sqrt(fan_in) if n != 768 else 0.35
        w[name] = torch.from_numpy(grid_ref.lcg_params(n, sc, 100 + k + seed_salt)).reshape(shp).to(dtype)
    return w


def golden_field(g, seed_salt, dtype=torch.float32, requires_grad=True):
    variant = str(g["variant"])
    specs = [grid_ref.make_grid_spec(int(g["log2_hash"][i]), float(g["per_level_scale"][i])) for i in range(2)]
    tabs = [torch.from_numpy(grid_ref.lcg_params(specs[i].n_params, 0.05, i + 1 + seed_salt)).to(dtype) for i in range(2)]
    w = golden_decoder_weights(variant, seed_salt, dtype)
    beta = torch.full((1,), 10.0, dtype=dtype)
    f = path_ref.Field(specs[0], specs[1], tabs[0], tabs[1], variant, w, beta, torch.from_numpy(g["bound"]).to(torch.float32))
    if requires_grad:
        for t in f.parameters():
            t.requires_grad_(True)
    return f


class DrawQueue:
    """Feeds recorded torch.rand draws to the oracle in reference order, checking shapes."""

    def __init__(self, draws):
        self.draws = list(draws)

    def __call__(self, shape):
        t = self.draws.pop(0)
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1); b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b, floor=1e-6):
    a = torch.as_tensor(a, dtype=torch.float64); b = torch.as_tensor(b, dtype=torch.float64)
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max())


def golden_img_draws(g):
    """The recorded torch.rand draws of a render_img fixture, in consumption order."""
    flat = torch.from_numpy(g["draws_flat"])
    out, o = [], 0
    for shp in g["draw_shapes"]:
        n = int(shp[0]) * int(shp[1])
        out.append(flat[o:o + n].reshape(int(shp[0]), int(shp[1]))); o += n
    assert o == flat.numel()
    return out


def img_draws_by_ray_slot(g):
    """Re-index the per-chunk compacted draws of a render_img fixture by ray slot, the layout the fused
    steps consume: t_rand (R,S), t_uni (R,n_strat), u_pdf (R,n_imp); rows of the other ray class stay 0."""
    ns, ni = int(g["n_stratified"]), int(g["n_importance"])
    S = ns + ni
    gt = torch.from_numpy(g["depth_img"]).reshape(-1)
    R, B = gt.numel(), int(g["ray_batch"])
    t_rand = torch.zeros((R, S)); t_uni = torch.zeros((R, ns)); u_pdf = torch.zeros((R, ni))
    draws = golden_img_draws(g)
    for i in range(0, R, B):
        m = gt[i:i + B] > 0
        rows = torch.arange(i, min(i + B, R))
        t_rand[rows[m]] = draws.pop(0)
        if int((~m).sum()) > 0:
            t_uni[rows[~m]] = draws.pop(0)
            u_pdf[rows[~m]] = draws.pop(0)
    assert not draws
    return t_rand, t_uni, u_pdf


# ---- host harness of the culling kernels' thread functions (tests/host_harness) ----
from host_harness.loader import cull_host, cull_host_compact, cull_host_frames, cull_host_hull, metrics_host  # noqa: E402,F401
