"""CPU: known-answer tests that pin the grid oracle (none exist upstream)."""
import numpy as np
import pytest
import torch

from oracle import grid_ref

# BASELINE.md section 2 tables
KAT = {
    (16, 816): dict(res=[16, 21, 28, 36, 46, 60, 78, 101, 131, 170, 221, 286, 372, 484, 628, 817], total=868400,
                    sizes=[4096, 9264, 21952, 46656] + [65536] * 12, hashed=[0] * 4 + [1] * 12),
    (19, 816): dict(res=[16, 21, 28, 36, 46, 60, 78, 101, 131, 170, 221, 286, 372, 484, 628, 817], total=5588448,
                    sizes=[4096, 9264, 21952, 46656, 97336, 216000, 474552] + [524288] * 9, hashed=[0] * 7 + [1] * 9),
    (16, 456): dict(res=[16, 21, 26, 32, 40, 49, 62, 77, 96, 120, 150, 187, 234, 292, 365, 456], total=848600,
                    sizes=[4096, 9264, 17576, 32768, 64000] + [65536] * 11, hashed=[0] * 5 + [1] * 11),
}


@pytest.mark.parametrize("key", list(KAT))
def test_level_tables_known_answers(key):
    log2T, res = key
    spec = grid_ref.make_grid_spec(log2T, grid_ref.per_level_scale_from_resolution(res))
    k = KAT[key]
    assert [l.res for l in spec.levels] == k["res"]
    assert [l.size for l in spec.levels] == k["sizes"]
    assert [int(l.hashed) for l in spec.levels] == k["hashed"]
    assert spec.total_entries == k["total"]
    assert [l.offset for l in spec.levels] == list(np.cumsum([0] + k["sizes"][:-1]))
    assert abs(grid_ref.per_level_scale_from_resolution(816) - 1.2996847159335432) < 1e-15
    assert abs(grid_ref.per_level_scale_from_resolution(456) - 1.2502292558175099) < 1e-15


def test_hash_known_answers():
    spec = grid_ref.make_grid_spec(16, grid_ref.per_level_scale_from_resolution(816))
    lv = 10
    assert spec.levels[lv].hashed
    assert grid_ref.grid_index(spec, lv, 1, 1, 1) == (1 ^ 2654435761 ^ 805459861) % 65536
    big = (3 ^ ((200 * 2654435761) & 0xFFFFFFFF) ^ ((150 * 805459861) & 0xFFFFFFFF)) % 65536      # uint32 wrap-around
    assert grid_ref.grid_index(spec, lv, 3, 200, 150) == big
    # dense level: linear index, and the x == 1 corner (g+1 == res) wraps through % size
    l0 = spec.levels[0]
    assert grid_ref.grid_index(spec, 0, 3, 2, 1) == 3 + 2 * 16 + 1 * 256
    assert grid_ref.grid_index(spec, 0, 16, 16, 16) == (16 + 16 * 16 + 16 * 256) % l0.size


def test_c_and_torch_restatements_agree_and_partition_unity():
    spec = grid_ref.make_grid_spec(16, grid_ref.per_level_scale_from_resolution(456))
    g = torch.Generator().manual_seed(0)
    x = torch.rand(3000, 3, generator=g)
    x[0] = 0.0; x[1] = 1.0
    for i, lv in enumerate(spec.levels):
        x[2 + i] = torch.tensor([(4 - 0.5) / lv.scale] * 3).clamp(0, 1)          # exactly on cell boundaries
    idx_c, w_c = grid_ref.c_corners(spec, x.numpy())
    idx_t, w_t = grid_ref.corner_tables(spec, x)
    assert np.array_equal(idx_c.astype(np.int64), idx_t.numpy())                   # bit-exact indices
    assert np.array_equal(w_c, w_t.numpy())                                        # identical fp32 weights
    assert np.abs(w_c.sum(-1) - 1).max() < 1e-6                                    # trilinear partition of unity
    params = grid_ref.lcg_params(spec.n_params, 0.05, 9)
    y_c = grid_ref.c_encode_fwd(spec, params, x.numpy())
    y_t = grid_ref.encode(spec, torch.from_numpy(params), x).numpy()
    assert np.abs(y_c - y_t).max() < 2e-7
    const = np.full(spec.n_params, 0.5, dtype=np.float32)
    assert np.abs(grid_ref.c_encode_fwd(spec, const, x.numpy()) - 0.5).max() < 1e-6
    # single-hot table entry -> output equals the exact interpolation weight
    hot = np.zeros(spec.n_params, dtype=np.float32)
    lv = spec.levels[7]
    e = int(idx_c[5, 7, 3])
    hot[(lv.offset + e) * 2] = 1.0
    y = grid_ref.c_encode_fwd(spec, hot, x[5:6].numpy())
    want = w_c[5, 7][idx_c[5, 7] == e].sum()
    assert abs(y[0, 14] - want) < 1e-7


def test_oracle_gradients_fp64():
    """gradcheck-style: analytic C backward (fp64 accumulation) vs autograd of the fp64 torch restatement."""
    spec = grid_ref.make_grid_spec(16, grid_ref.per_level_scale_from_resolution(456))
    g = torch.Generator().manual_seed(1)
    x = torch.rand(500, 3, generator=g)
    params = torch.from_numpy(grid_ref.lcg_params(spec.n_params, 0.05, 2))
    dy = torch.randn(500, 32, generator=g)
    p64 = params.double().requires_grad_(True)
    x64 = x.double().requires_grad_(True)
    y = grid_ref.encode(spec, p64, x64)
    y.backward(dy.double())
    gp = grid_ref.c_encode_bwd_params(spec, x.numpy(), dy.numpy())
    # fp32 interpolation weights (w = pos - floor(pos) with pos up to ~800 carries ~5e-5 absolute error) vs fp64 weights
    assert np.linalg.norm(gp - p64.grad.numpy()) / np.linalg.norm(gp) < 1e-4
    gx = grid_ref.c_encode_bwd_input(spec, params.numpy(), x.numpy(), dy.numpy())
    assert np.linalg.norm(gx - x64.grad.numpy()) / np.linalg.norm(gx) < 1e-4


def test_compositing_gradcheck_fp64():
    from oracle import path_ref
    g = torch.Generator().manual_seed(3)
    raw = torch.rand(5, 12, 4, generator=g, dtype=torch.float64)
    raw[..., 3] = raw[..., 3] * 0.4 - 0.2
    raw.requires_grad_(True)
    z = torch.sort(torch.rand(5, 12, generator=g, dtype=torch.float64) * 3, -1)[0]
    beta = torch.tensor([10.0], dtype=torch.float64, requires_grad=True)
    fn = lambda r, b: torch.cat([t.reshape(5, -1) for t in path_ref.composite(r, z, b)[:4]], -1)
    assert torch.autograd.gradcheck(fn, (raw, beta), eps=1e-6, atol=1e-5)
