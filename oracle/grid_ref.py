"""oracle/grid_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

tiny-cuda-nn HashGrid encoding restated as differentiable PyTorch ops (CPU, fp32 or fp64),
plus numpy front-ends to the plain-C restatement in grid_oracle.c.

PARITY UNPINNED for the tcnn arithmetic (source absent; see oracle/__init__.py): the
algorithm follows SURVEY.md section 8a-1; the level tables are pinned by BASELINE.md's KATs.

Reference call sites: src/UNISLAM.py:224-259 (construction), src/networks/decoders.py:101-103
(forward), src/Mapper.py:118-121 (.params into Adam).
"""
import ctypes
import math
from dataclasses import dataclass
from typing import List

import numpy as np
import torch

from . import OrcLevel, lib

PRIME_Y = 2654435761
PRIME_Z = 805459861
U32 = 0xFFFFFFFF
N_FEATS = 2


@dataclass
class Level:
    scale: float      # fp32 value held in a python float
    res: int
    size: int
    offset: int
    hashed: bool


@dataclass
class GridSpec:
    n_levels: int
    log2_hashmap_size: int
    base_resolution: int
    per_level_scale: float
    levels: List[Level]
    total_entries: int

    @property
    def n_params(self) -> int:
        return self.total_entries * N_FEATS

    @property
    def n_output_dims(self) -> int:
        return self.n_levels * N_FEATS

    def c_levels(self):
        arr = (OrcLevel * self.n_levels)()
        for i, lv in enumerate(self.levels):
            arr[i] = OrcLevel(lv.scale, lv.res, lv.size, lv.offset, int(lv.hashed))
        return arr


def per_level_scale_from_resolution(desired_resolution: int, n_levels: int = 16) -> float:
    """src/UNISLAM.py:241 (divides by n_levels where base_resolution is meant; both are 16)."""
    return float(np.exp2(np.log2(desired_resolution / n_levels) / (n_levels - 1)))


def make_grid_spec(log2_hashmap_size: int, per_level_scale: float, n_levels: int = 16,
                   base_resolution: int = 16) -> GridSpec:
    arr = (OrcLevel * n_levels)()
    total = ctypes.c_uint32(0)
    rc = lib().orc_grid_levels(n_levels, log2_hashmap_size, base_resolution, float(per_level_scale),
                               arr, ctypes.byref(total))
    assert rc == 0
    levels = [Level(float(a.scale), int(a.res), int(a.size), int(a.offset), bool(a.hashed)) for a in arr]
    return GridSpec(n_levels, log2_hashmap_size, base_resolution, float(per_level_scale), levels, int(total.value))


def init_params(spec: GridSpec, seed: int = 0, kind: str = "tcnn") -> torch.Tensor:
    """tcnn init U(-1e-4, 1e-4) or the 'trained-like' N(0, 0.05) set."""
    g = torch.Generator().manual_seed(seed)
    if kind == "tcnn":
        return (torch.rand(spec.n_params, generator=g) * 2 - 1) * 1e-4
    if kind == "trained":
        return torch.randn(spec.n_params, generator=g) * 0.05
    raise ValueError(kind)


def lcg_params(n: int, scale: float = 0.1, salt: int = 0) -> np.ndarray:
    """Platform-independent pseudo-random fp32 fill (pure integer hashing) for golden fixtures."""
    with np.errstate(over="ignore"):
        i = np.arange(n, dtype=np.uint64) + np.uint64((salt * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
        i = (i ^ (i >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        i = (i ^ (i >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        i = i ^ (i >> np.uint64(31))
    u = (i >> np.uint64(40)).astype(np.float64) / float(1 << 24)      # 24-bit uniform in [0,1)
    return ((u - 0.5) * 2.0 * scale).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# torch restatement (differentiable wrt params and x)
# ----------------------------------------------------------------------------------------------
def _level_index(lv: Level, gx, gy, gz):
    """tcnn grid_index<3,CoherentPrime> on int64 tensors holding uint32 values."""
    if lv.hashed:
        idx = (gx & U32) ^ ((gy * PRIME_Y) & U32) ^ ((gz * PRIME_Z) & U32)
    else:
        idx = (gx + gy * lv.res + gz * (lv.res * lv.res)) & U32
    return idx % lv.size


def _pos_fract(lv: Level, x: torch.Tensor):
    """pos = fmaf(scale, x, .5) in fp32 (double product+add rounded once); g = floor; w = pos-g.
    Returns int64 g (N,3) and w (N,3) in x.dtype, w differentiable wrt x (dw/dx = scale)."""
    x32 = x.detach().to(torch.float32)
    pos32 = (x32.double() * lv.scale + 0.5).to(torch.float32)
    g = torch.floor(pos32)
    if x.dtype == torch.float32:
        # straight-through construction keeps the exact fp32 values and d w/d x = scale
        w_val = pos32 - g
        w = w_val + (x - x.detach()) * lv.scale if x.requires_grad else w_val
    else:
        w = (x * lv.scale + 0.5) - g.to(x.dtype)
    return g.to(torch.int64), w


def corner_tables(spec: GridSpec, x: torch.Tensor):
    """Indices (N,L,8) int64 within-level and weights (N,L,8) for every corner."""
    idxs, ws = [], []
    for lv in spec.levels:
        g, w = _pos_fract(lv, x)
        ci, cw = [], []
        for c in range(8):
            wt = None
            gl = []
            for d in range(3):
                if c & (1 << d):
                    f = w[:, d]; gl.append(g[:, d] + 1)
                else:
                    f = 1.0 - w[:, d]; gl.append(g[:, d])
                wt = f if wt is None else wt * f
            ci.append(_level_index(lv, gl[0], gl[1], gl[2]))
            cw.append(wt)
        idxs.append(torch.stack(ci, -1))
        ws.append(torch.stack(cw, -1))
    return torch.stack(idxs, 1), torch.stack(ws, 1)


def encode(spec: GridSpec, params: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """y (N, 2L) = HashGrid(x). x in [0,1]^3. dtype follows params (fp32 oracle / fp64 check)."""
    table = params.reshape(-1, N_FEATS)
    xx = x.to(params.dtype) if x.dtype != params.dtype else x
    outs = []
    for lv in spec.levels:
        g, w = _pos_fract(lv, xx)
        acc = None
        for c in range(8):
            wt = None
            gl = []
            for d in range(3):
                if c & (1 << d):
                    f = w[:, d]; gl.append(g[:, d] + 1)
                else:
                    f = 1.0 - w[:, d]; gl.append(g[:, d])
                wt = f if wt is None else wt * f
            e = _level_index(lv, gl[0], gl[1], gl[2]) + lv.offset
            v = table[e]
            term = wt[:, None] * v
            acc = term if acc is None else acc + term
        outs.append(acc)
    return torch.cat(outs, -1)


# ----------------------------------------------------------------------------------------------
# numpy front-ends to the C restatement
# ----------------------------------------------------------------------------------------------
def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def c_corners(spec: GridSpec, x: np.ndarray):
    x = _f32(x); n = x.shape[0]
    idx = np.empty((n, spec.n_levels, 8), dtype=np.uint32)
    w = np.empty((n, spec.n_levels, 8), dtype=np.float32)
    lib().orc_grid_corners(spec.c_levels(), spec.n_levels, x.ctypes.data, n, idx.ctypes.data, w.ctypes.data)
    return idx, w


def c_encode_fwd(spec: GridSpec, params: np.ndarray, x: np.ndarray) -> np.ndarray:
    x = _f32(x); params = _f32(params); n = x.shape[0]
    y = np.empty((n, spec.n_output_dims), dtype=np.float32)
    lib().orc_grid_encode_fwd(spec.c_levels(), spec.n_levels, params.ctypes.data, x.ctypes.data, n, y.ctypes.data)
    return y


def c_encode_bwd_params(spec: GridSpec, x: np.ndarray, dy: np.ndarray) -> np.ndarray:
    x = _f32(x); dy = _f32(dy); n = x.shape[0]
    grad = np.zeros(spec.n_params, dtype=np.float64)
    lib().orc_grid_encode_bwd_params(spec.c_levels(), spec.n_levels, x.ctypes.data, dy.ctypes.data, n, grad.ctypes.data)
    return grad


def c_encode_bwd_input(spec: GridSpec, params: np.ndarray, x: np.ndarray, dy: np.ndarray) -> np.ndarray:
    x = _f32(x); dy = _f32(dy); params = _f32(params); n = x.shape[0]
    dx = np.empty((n, 3), dtype=np.float32)
    lib().orc_grid_encode_bwd_input(spec.c_levels(), spec.n_levels, params.ctypes.data, x.ctypes.data,
                                    dy.ctypes.data, n, dx.ctypes.data)
    return dx


def grid_index(spec: GridSpec, level: int, gx: int, gy: int, gz: int) -> int:
    arr = spec.c_levels()
    return int(lib().orc_grid_index(ctypes.byref(arr[level]), gx & U32, gy & U32, gz & U32))
