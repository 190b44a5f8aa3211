"""-m gpu parity tests: the CUDA path (through the C-ABI) against the oracle and the golden fixtures.

Bars (BASELINE.json north_star): hash / sample indices bit-exact; rendered depth / colour and losses
<= 1e-4 relative (fp32); parameter and pose gradients <= 1e-3 relative.
"""
import copy
import pickle

import numpy as np
import pytest
import torch

from oracle import grid_ref, path_ref
from helpers import DrawQueue, golden_field, golden_img_draws, img_draws_by_ray_slot, load_golden, max_rel, pkg, rel_err
import gpu_cases

pytestmark = pytest.mark.gpu
T = torch.from_numpy
DEV = "cuda:0"

GRIDS = {"replica_sdf": (16, 816), "replica_rgb": (19, 816), "scannet": (16, 456)}


def _points(n, spec, seed=0):
    """Random points plus the edge cases the index math must survive: 0, 1, cell boundaries."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, generator=g)
    x[0] = 0.0; x[1] = 1.0; x[2] = torch.tensor([0.0, 1.0, 0.5])
    for i, lv in enumerate(spec.levels):          # x = (k - 0.5)/scale sits exactly on a cell boundary
        x[3 + i] = torch.tensor([(3 - 0.5) / lv.scale, (5 - 0.5) / lv.scale, (7 - 0.5) / lv.scale]).clamp(0, 1)
    return x


@pytest.mark.parametrize("name", list(GRIDS))
def test_grid_indices_bit_exact_and_features(name):
    P = pkg()
    log2T, res = GRIDS[name]
    pls = grid_ref.per_level_scale_from_resolution(res)
    spec = grid_ref.make_grid_spec(log2T, pls)
    enc = P.Encoding(3, {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": log2T,
                         "base_resolution": 16, "per_level_scale": pls}, dtype=torch.float).to(DEV)
    assert [(round(a[0], 6), a[1], a[2], a[3], a[4]) for a in enc.level_table()] == \
           [(round(l.scale, 6), l.res, l.size, l.offset, l.hashed) for l in spec.levels]
    params = grid_ref.lcg_params(spec.n_params, 0.05, 3)
    with torch.no_grad():
        enc.params.copy_(T(params))
    x = _points(20000, spec)
    idx_ref, _ = grid_ref.c_corners(spec, x.numpy())
    idx = P.ops.grid_corner_indices(x.to(DEV), enc.grid).cpu().numpy()
    assert np.array_equal(idx.astype(np.uint32), idx_ref)                       # bit-exact hash / table indices
    y_ref = grid_ref.c_encode_fwd(spec, params, x.numpy())
    y = enc(x.to(DEV)).detach().cpu().numpy()
    assert np.abs(y - y_ref).max() <= 1e-6


@pytest.mark.parametrize("name", ["replica_sdf", "scannet"])
def test_grid_backward(name):
    P = pkg()
    log2T, res = GRIDS[name]
    pls = grid_ref.per_level_scale_from_resolution(res)
    spec = grid_ref.make_grid_spec(log2T, pls)
    enc = P.Encoding(3, {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": log2T,
                         "base_resolution": 16, "per_level_scale": pls}, dtype=torch.float).to(DEV)
    params = grid_ref.lcg_params(spec.n_params, 0.05, 4)
    with torch.no_grad():
        enc.params.copy_(T(params))
    x = _points(8192, spec, 1)
    dy = torch.randn(8192, 32, generator=torch.Generator().manual_seed(2))
    xg = x.to(DEV).requires_grad_(True)
    y = enc(xg)
    y.backward(dy.to(DEV))
    gp_ref = grid_ref.c_encode_bwd_params(spec, x.numpy(), dy.numpy())            # fp64 accumulation
    gp = enc.params.grad.cpu().double().numpy()
    assert np.linalg.norm(gp - gp_ref) / np.linalg.norm(gp_ref) < 1e-5
    touched = gp_ref != 0
    assert np.array_equal(gp != 0, touched) or (np.abs(gp[~touched]).max() == 0)
    assert (np.abs(gp - gp_ref)[touched] / np.maximum(np.abs(gp_ref[touched]), 1e-3)).max() < 1e-3
    gx_ref = grid_ref.c_encode_bwd_input(spec, params, x.numpy(), dy.numpy())
    gx = xg.grad.cpu().numpy()
    assert np.linalg.norm(gx - gx_ref) / np.linalg.norm(gx_ref) < 1e-4


def test_grid_properties_full_size():
    """Size-independent properties at a BASELINE-sized batch (240k points, Replica colour grid)."""
    P = pkg()
    pls = grid_ref.per_level_scale_from_resolution(816)
    enc = P.Encoding(3, {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19,
                         "base_resolution": 16, "per_level_scale": pls}, dtype=torch.float).to(DEV)
    n = 6000 * 40
    x = torch.rand(n, 3, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
    with torch.no_grad():
        enc.params.fill_(0.25)                      # constant table -> constant output (partition of unity)
    y = enc(x)
    assert (y - 0.25).abs().max() < 1e-6
    dy = torch.randn(n, 32, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    enc.params.grad = None
    y.backward(dy)
    g = enc.params.grad.view(-1, 2).double()
    # weights of the 8 corners sum to 1 => per level, sum of scattered gradient == sum of dy
    off = 0
    for l, (_, _, size, offset, _) in enumerate(enc.level_table()):
        got = g[offset:offset + size].sum(0)
        want = dy[:, 2 * l:2 * l + 2].double().sum(0)
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-2), l
    # linearity: encode(a*T1 + b*T2) == a*encode(T1) + b*encode(T2)
    t1 = torch.randn_like(enc.params) * 0.05; t2 = torch.randn_like(enc.params) * 0.05
    with torch.no_grad():
        enc.params.copy_(t1); y1 = enc(x[:50000]).clone()
        enc.params.copy_(t2); y2 = enc(x[:50000]).clone()
        enc.params.copy_(0.5 * t1 - 2.0 * t2); y3 = enc(x[:50000])
    assert (y3 - (0.5 * y1 - 2.0 * y2)).abs().max() < 1e-5


def test_encoding_module_contract():
    """B1 seam: flat leaf Parameter, deepcopy (Tracker.py:105-108), pickle (mp.spawn, UNISLAM.py:295-298), empty batch."""
    P = pkg()
    enc = P.Encoding(3, {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 16,
                         "base_resolution": 16, "per_level_scale": 1.3}, dtype=torch.float).to(DEV)
    assert enc.params.is_leaf and enc.params.dim() == 1 and enc.n_output_dims == 32
    assert float(enc.params.abs().max()) <= 1e-4
    x = torch.rand(1000, 3, device=DEV)
    y = enc(x)
    e2 = copy.deepcopy(enc)
    e3 = pickle.loads(pickle.dumps(enc))
    assert torch.equal(e2(x), y) and torch.equal(e3(x), y)
    assert enc(torch.empty(0, 3, device=DEV)).shape == (0, 32)
    with pytest.raises(RuntimeError):
        enc(torch.rand(4, 3))                        # CPU tensor: no fallback


@pytest.mark.parametrize("variant", ["A", "B"])
def test_network_matches_torch(variant):
    P = pkg()
    n = 5000
    g = torch.Generator().manual_seed(0)
    h = torch.randn(n, 32, generator=g)
    if variant == "B":
        for n_out, act, tact in ((1, "Tanh", torch.tanh), (3, "Sigmoid", torch.sigmoid)):
            net = P.Network(32, n_out, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": act,
                                        "n_neurons": 16, "n_hidden_layers": 1}).to(DEV)
            hp = h.to(DEV).requires_grad_(True)
            out = net(hp)
            dout = torch.randn(n, n_out, generator=g)
            out.backward(dout.to(DEV))
            p = net.params.detach().cpu().double().requires_grad_(True)
            hr = h.double().requires_grad_(True)
            ref = path_ref.mlp_B(hr, p, n_out, tact)
            ref.backward(dout.double())
            assert (out.detach().cpu().double() - ref.detach()).abs().max() < 1e-5      # O(1) activations: absolute fp32 noise
            assert rel_err(net.params.grad.cpu(), p.grad) < 1e-4
            assert rel_err(hp.grad.cpu(), hr.grad) < 1e-4
    else:
        # variant A (tcnn_network: False): the nn.Linear stacks of decoders.py:72-84,125-128,150-153 through the same
        # stand-alone decoder seam (usl_mlp_fwd / usl_mlp_bwd), against torch.nn.functional in fp64
        import torch.nn.functional as Fn
        from importlib import import_module
        modules = import_module("uni-slam_b200.modules")
        for n_out, act_id, tact in ((1, P._lib.ACT_TANH, torch.tanh), (3, P._lib.ACT_SIGMOID, torch.sigmoid)):
            lin = lambda o, i: [((torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5), ((torch.rand(o, generator=g) * 2 - 1) / i ** 0.5)]
            ws = lin(16, 32) + lin(16, 16) + lin(n_out, 16)
            wd = [w.to(DEV).requires_grad_(True) for w in ws]
            hp = h.to(DEV).requires_grad_(True)
            out = modules._MlpFn.apply(P.ops.DecoderLayout("A", n_out, act_id), hp, *wd)
            dout = torch.randn(n, n_out, generator=g)
            out.backward(dout.to(DEV))
            wr = [w.double().requires_grad_(True) for w in ws]
            hr = h.double().requires_grad_(True)
            a = torch.relu(Fn.linear(hr, wr[0], wr[1]))
            a = torch.relu(Fn.linear(a, wr[2], wr[3]))
            ref = tact(Fn.linear(a, wr[4], wr[5]))
            ref.backward(dout.double())
            assert (out.detach().cpu().double() - ref.detach()).abs().max() < 1e-5
            for got, want in zip(wd, wr):
                assert rel_err(got.grad.cpu(), want.grad) < 1e-4
            assert rel_err(hp.grad.cpu(), hr.grad) < 1e-4


@pytest.mark.parametrize("name", ["map_replica_k1", "map_replica_k7", "map_scannet_k23", "map_replica_nomask", "map_replica_kfstore"])
def test_mapping_step_matches_reference(name):
    r = gpu_cases.run_mapping_case(name, DEV)
    print(name, r)
    assert r["rays_o_mismatch"] == 0 and r["rays_d_mismatch"] == 0 and r["valid_mismatch"] == 0
    assert r["z_depth_mismatch"] == 0                                  # sample positions bit-exact
    assert r["z_hole_maxabs"] < 1e-4
    assert r["pdf_inds_mismatch"] == 0                                 # searchsorted indices of sample_pdf bit-exact
    for k in ("term_rel", "pixel_unc_rel", "depth_rel", "rgb_rel", "loss_rel"):
        assert r[k] < 1e-4, (k, r[k])
    for k in ("dec_grad_rel", "beta_grad_rel", "table_grad_rel", "pose_grad_rel"):
        assert r[k] < 1e-3, (k, r[k])
    for pre in ("grad_sdf_table", "grad_rgb_table"):
        assert r[pre + "_support_miss"] == 0 and r[pre + "_nnz_excess"] == 0, pre


@pytest.mark.parametrize("name", ["track_replica", "track_scannet", "track_scannet_nomask"])
def test_tracking_step_matches_reference(name):
    r = gpu_cases.run_tracking_case(name, DEV)
    print(name, r)
    assert r["rays_o_mismatch"] == 0 and r["rays_d_mismatch"] == 0 and r["valid_mismatch"] == 0 and r["z_mismatch"] == 0
    for k in ("term_rel", "pixel_unc_rel", "depth_rel", "rgb_rel", "loss_rel"):
        assert r[k] < 1e-4, (k, r[k])
    assert r["mean_punc_err"] < 1e-3
    assert r["grad_T_rel"] < 1e-3 and r["grad_R_rel"] < 1e-3


@pytest.mark.parametrize("name", ["map_replica_k7", "map_scannet_k23"])
def test_dropin_renderer_matches_reference(name, monkeypatch):
    """B3/B4 seams: modules.Decoders + modules.Renderer.render_batch_ray fed the reference's own draws,
    with the loss assembled by the oracle's restatement of the host code (torch ops on CUDA)."""
    P = pkg()
    g = load_golden(name)
    variant = str(g["variant"])
    cfg = {"grid_mode": "hash_grid", "grid": {"tcnn_network": variant == "B"}, "scale": 1,
           "rendering": {"perturb": True, "n_stratified": int(g["n_stratified"]), "n_importance": int(g["n_importance"])}}
    meta, tabs, dec, beta = gpu_cases.cuda_field(g, 0, DEV)
    encs = []
    for i in range(2):
        e = P.Encoding(3, {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": int(g["log2_hash"][i]),
                           "base_resolution": 16, "per_level_scale": float(g["per_level_scale"][i])}, dtype=torch.float).to(DEV)
        with torch.no_grad():
            e.params.copy_(tabs[i])
        encs.append(e)
    decoders = P.Decoders(cfg, c_dim=32, truncation=float(g["truncation"]), learnable_beta=True).to(DEV)
    with torch.no_grad():
        for t, src in zip(decoders.decoder_tensors(), dec):
            t.copy_(src)
    H, W, fx, fy, cx, cy = [float(v) for v in g["meta_H_W_fx_fy_cx_cy"]]
    fake = type("U", (), dict(bound=T(g["bound"]), device=DEV, H=int(H), W=int(W), fx=fx, fy=fy, cx=cx, cy=cy))()
    renderer = P.Renderer(cfg, fake)
    draws = [T(g["t_rand"]).to(DEV)]
    if "t_rand_uni" in g:
        draws += [T(g["t_rand_uni"]).to(DEV), T(g["u_pdf"]).to(DEV)]
    q = DrawQueue(draws)
    monkeypatch.setattr(torch, "rand", lambda shape, device=None, **k: q(tuple(shape)))
    rays_o = T(g["render_rays_o"]).to(DEV).requires_grad_(True)
    rays_d = T(g["render_rays_d"]).to(DEV).requires_grad_(True)
    gt_depth = T(g["render_gt_depth"]).to(DEV)
    scene_rep = ([encs[0]], [encs[1]])
    ret = renderer.render_batch_ray(scene_rep, decoders, rays_d, rays_o, DEV, float(g["truncation"]), gt_depth=gt_depth)
    monkeypatch.undo()
    for nm, t in zip(("term", "pixel_unc", "depth", "rgb", "sdf"), ret[:5]):
        assert max_rel(t.detach().cpu(), g["ret_" + nm], 1e-3 if nm != "sdf" else 1e-2) < 1e-4, nm
    has = gt_depth.cpu() > 0
    assert torch.equal(ret[5].cpu()[has], T(g["ret_z_vals"])[has])
    # the reference's own loss code (restated) on top of our outputs, then autograd through our backward
    gt_color = torch.cat([T(g[f"call{ci}_out_color"]) for ci in range(int(g["n_sample_calls"]))])
    ro_all = torch.cat([T(g[f"call{ci}_out_rays_o"]) for ci in range(int(g["n_sample_calls"]))])
    rd_all = torch.cat([T(g[f"call{ci}_out_rays_d"]) for ci in range(int(g["n_sample_calls"]))])
    d_all = torch.cat([T(g[f"call{ci}_out_depth"]) for ci in range(int(g["n_sample_calls"]))])
    inside = path_ref.bbox_exit(ro_all, rd_all, T(g["bound"])) >= d_all
    loss = path_ref.mapping_loss(ret, gt_depth, gt_color[inside].to(DEV), float(g["truncation"]))
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-4
    loss.backward()
    names = gpu_cases.DEC_ORDER[variant]
    for nm, t in zip(names, decoders.decoder_tensors()):
        assert rel_err(t.grad.cpu(), g["grad_dec." + nm]) < 1e-3, nm
    assert rel_err(decoders.beta.grad.cpu(), g["grad_dec.beta"]) < 1e-3
    for pre, e in (("grad_sdf_table", encs[0]), ("grad_rgb_table", encs[1])):
        assert rel_err(e.params.grad.cpu()[T(g[pre + "_idx"])], g[pre + "_val"]) < 1e-3
    assert rays_o.grad is not None and torch.isfinite(rays_o.grad).all() and torch.isfinite(rays_d.grad).all()
    # ray gradients numerically: the oracle's autograd through the same rays / draws (this is what feeds the pose gradient)
    field = golden_field(g, 0)
    ro = T(g["render_rays_o"]).clone().requires_grad_(True); rd = T(g["render_rays_d"]).clone().requires_grad_(True)
    qd = [T(g["t_rand"])] + ([T(g["t_rand_uni"]), T(g["u_pdf"])] if "t_rand_uni" in g else [])
    ret_ref = path_ref.render_batch_ray(field, rd, ro, T(g["render_gt_depth"]), int(g["n_stratified"]), int(g["n_importance"]),
                                        float(g["truncation"]), *qd)
    path_ref.mapping_loss(ret_ref, T(g["render_gt_depth"]), gt_color[inside], float(g["truncation"])).backward()
    assert rel_err(rays_o.grad.cpu(), ro.grad) < 1e-3 and rel_err(rays_d.grad.cpu(), rd.grad) < 1e-3


def _dropin_modules(P, g, seed_salt):
    variant = str(g["variant"])
    cfg = {"grid_mode": "hash_grid", "grid": {"tcnn_network": variant == "B"}, "scale": 1,
           "rendering": {"perturb": True, "n_stratified": int(g["n_stratified"]), "n_importance": int(g["n_importance"])}}
    meta, tabs, dec, beta = gpu_cases.cuda_field(g, seed_salt, DEV)
    encs = []
    for i in range(2):
        e = P.Encoding(3, {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": int(g["log2_hash"][i]),
                           "base_resolution": 16, "per_level_scale": float(g["per_level_scale"][i])}, dtype=torch.float).to(DEV)
        with torch.no_grad():
            e.params.copy_(tabs[i])
        encs.append(e)
    decoders = P.Decoders(cfg, c_dim=32, truncation=float(g["truncation"]), learnable_beta=True).to(DEV)
    with torch.no_grad():
        for t, src in zip(decoders.decoder_tensors(), dec):
            t.copy_(src)
    return cfg, encs, decoders


def _check_img(ret, g, where):
    """depth / colour / termination / uncertainties of a rendered frame vs the reference's render_img output."""
    for nm, t in zip(("depth", "color", "term", "pixel_unc", "depth_unc"), ret):
        ref = T(g["ret_" + nm]).double().reshape(-1)
        got = t.detach().cpu().double().reshape(-1)
        # 1e-4 relative on the rendered values (north-star tolerance); the two uncertainties are differences of nearly
        # equal numbers (1 - sum w, depth - z), so they get the same tolerance on an absolute floor of their scale
        floor = 1e-3 if nm in ("pixel_unc", "depth_unc") else 1e-4
        assert max_rel(got, ref, floor) < 1e-4, (where, nm)


@pytest.mark.parametrize("name", ["img_replica", "img_scannet"])
def test_dropin_render_img_matches_reference(name, monkeypatch):
    """f3 / B4: modules.Renderer.render_img (whole frame in ray_batch_size chunks, reference RNG order) vs the
    unmodified reference's Renderer.render_img output (tests/golden/img_*.npz)."""
    P = pkg()
    g = load_golden(name)
    cfg, encs, decoders = _dropin_modules(P, g, 80)
    H, W, fx, fy, cx, cy = [float(v) for v in g["meta_H_W_fx_fy_cx_cy"]]
    fake = type("U", (), dict(bound=T(g["bound"]), device=DEV, H=int(H), W=int(W), fx=fx, fy=fy, cx=cx, cy=cy))()
    renderer = P.Renderer(cfg, fake, ray_batch_size=int(g["ray_batch"]))
    q = DrawQueue([d.to(DEV) for d in golden_img_draws(g)])
    monkeypatch.setattr(torch, "rand", lambda shape, device=None, **k: q(tuple(shape)))
    ret = renderer.render_img(([encs[0]], [encs[1]]), decoders, T(g["c2w"]), float(g["truncation"]), DEV, gt_depth=T(g["depth_img"]).to(DEV))
    monkeypatch.undo()
    assert not q.draws                                                            # consumed exactly the reference's draws
    for nm, t in zip(("depth", "color", "term", "pixel_unc", "depth_unc"), ret):
        assert str(t.dtype) == str(g["dtype_" + nm]) and tuple(t.shape) == g["ret_" + nm].shape, nm
    _check_img(ret, g, "drop-in")


@pytest.mark.parametrize("name", ["img_replica", "img_scannet"])
def test_render_image_step_matches_reference(name):
    """f3: the fused forward-only frame renderer (in-kernel get_rays, masks instead of compaction) vs the reference's
    render_img output; ragged chunks, and a row-sharded render (two pixel ranges) must give the same frame."""
    P = pkg()
    g = load_golden(name)
    meta, tabs, dec, beta = gpu_cases.cuda_field(g, 80, DEV)
    H, W, fx, fy, cx, cy = [float(v) for v in g["meta_H_W_fx_fy_cx_cy"]]
    H, W = int(H), int(W)
    t_rand, t_uni, u_pdf = [t.to(DEV) for t in img_draws_by_ray_slot(g)]
    c2w, dep = T(g["c2w"]).to(DEV), T(g["depth_img"]).to(DEV)
    kw = dict(n_stratified=int(g["n_stratified"]), n_importance=int(g["n_importance"]), truncation=float(g["truncation"]),
              H=H, W=W, fx=fx, fy=fy, cx=cx, cy=cy)
    keys = ("depth", "color", "term", "pixel_unc", "depth_unc")
    whole = P.RenderImageStep(meta, tabs[0], tabs[1], dec, beta, chunk_rays=H * W, **kw).run(c2w, dep, t_rand, t_uni, u_pdf)
    _check_img([whole[k] for k in keys], g, "one chunk")
    step = P.RenderImageStep(meta, tabs[0], tabs[1], dec, beta, chunk_rays=257, **kw)      # ragged chunks
    chunked = step.run(c2w, dep, t_rand, t_uni, u_pdf)
    for k in keys:
        assert torch.equal(chunked[k], whole[k]), k                                # chunking must not change a bit
    cut = (H // 2) * W                                                             # row shard, as parallel.slab_range cuts it
    a = step.run(c2w, dep, t_rand[:cut], t_uni[:cut], u_pdf[:cut], pixel_begin=0, pixel_end=cut)
    b = step.run(c2w, dep, t_rand[cut:], t_uni[cut:], u_pdf[cut:], pixel_begin=cut, pixel_end=H * W)
    for k in keys:
        assert torch.equal(torch.cat([a[k], b[k]]), whole[k]), k


def test_dense_sdf_query_matches_oracle():
    """a-11: in-kernel point generation + SDF-only field query vs the oracle's eval_points, slab-wise."""
    P = pkg()
    g = load_golden("map_replica_k7")
    meta, tabs, dec, beta = gpu_cases.cuda_field(g, 0, DEV)
    field = golden_field(g, 0, requires_grad=False)
    mc_bound = [[-1.0, 7.0], [-1.3, 3.7], [-1.7, 1.4]]
    axes = path_ref.mesh_grid_axes(mc_bound, resolution=0.25)            # coarse grid: seconds on the CPU oracle
    pts = path_ref.mesh_grid_points(axes)
    with torch.no_grad():
        ref = path_ref.eval_points_sdf(field, pts)
    q = P.DenseSdfQuery(meta, tabs[0], tabs[1], dec, [a.to(DEV) for a in axes])
    ny = q.ny
    parts = [q.run(0, ny // 3), q.run(ny // 3, ny)]                      # two slabs == one full query
    out = torch.cat(parts).cpu()
    assert out.shape == ref.shape
    assert torch.equal(out == -1, ref == -1)                              # strict in-bound mask identical
    assert (out - ref).abs().max() < 2e-5


def test_dense_sdf_query_fine_grid_matches_oracle():
    """The production spacing (1 cm): neighbouring lanes sit in the same or adjacent columns on EVERY level, i.e. the x-line
    decode (one collapsed column per lane + a shuffle) runs on all 16 levels, incl. ragged row ends, the far bound (x = 1
    wrap) and points outside the open bound.  Against the oracle's eval_points on the same points."""
    P = pkg()
    g = load_golden("map_replica_k7")
    meta, tabs, dec, beta = gpu_cases.cuda_field(g, 0, DEV)
    field = golden_field(g, 0, requires_grad=False)
    b = T(g["bound"])
    ax = [torch.arange(float(b[0, 1]) - 0.835, float(b[0, 1]) + 0.06, 0.01),        # 90 points along x, crossing the far bound
          torch.arange(float(b[1, 0]) - 0.02, float(b[1, 0]) + 0.05, 0.01),         # 7 rows, the first two outside
          torch.arange(0.1, 0.23, 0.01)]                                            # 13 z: one full and one ragged z tile
    pts = path_ref.mesh_grid_points(ax)
    with torch.no_grad():
        ref = path_ref.eval_points_sdf(field, pts)
    q = P.DenseSdfQuery(meta, tabs[0], tabs[1], dec, [a.to(DEV) for a in ax])
    got = q.run(0, q.ny).cpu()
    assert torch.equal(got == -1, ref == -1) and int((ref == -1).sum()) > 0
    assert (got - ref).abs().max() < 2e-6


@pytest.mark.parametrize("name", ["mesh_replica", "mesh_scannet"])
def test_dense_sdf_query_matches_reference(name):
    """a-11 against the unmodified reference: Mesher.get_grid_uniform + eval_points output (tests/golden/mesh_*.npz),
    both decoder variants, ragged tile edges of the x-tiled kernel, slabs of uneven height."""
    P = pkg()
    g = load_golden(name)
    meta, tabs, dec, beta = gpu_cases.cuda_field(g, 120, DEV)
    axes = [T(g["axis_" + nm]).float().to(DEV) for nm in "xyz"]
    q = P.DenseSdfQuery(meta, tabs[0], tabs[1], dec, axes)
    ny = q.ny
    cuts = [0, 1, ny // 2, ny]
    out = torch.cat([q.run(a, b) for a, b in zip(cuts[:-1], cuts[1:])]).cpu()
    ref = T(g["sdf"])
    assert out.shape == ref.shape
    assert torch.equal(out == -1, ref == -1)                              # strict in-bound mask identical
    assert (out - ref).abs().max() < 2e-5


def test_fused_adam_matches_torch_adam():
    """a-12 / f1: usl_adam_step vs torch.optim.Adam with the host code's group structure (Mapper.py:111-139; Tracker betas)."""
    P = pkg()
    g = torch.Generator().manual_seed(0)
    shapes = [(768,), (100003,), (16, 32), (1,), (20, 7)]
    ref = [torch.randn(s, generator=g).to(DEV).requires_grad_(True) for s in shapes]
    ours = [r.detach().clone().requires_grad_(True) for r in ref]
    mk = lambda ps: [{"params": [ps[0], ps[2], ps[3]], "lr": 1e-3}, {"params": [ps[1]], "lr": 0.05},
                     {"params": [ps[4]], "lr": 2e-3, "betas": (0.5, 0.999)}]
    o_ref = torch.optim.Adam(mk(ref))
    o_our = P.FusedAdam(mk(ours))
    for it in range(5):
        for r, o in zip(ref, ours):
            gr = torch.randn(r.shape, generator=g).to(DEV) * (10.0 ** (it - 2))
            gr[gr.abs() < 0.3 * gr.abs().max()] = 0                     # sparse gradients like the hash tables'
            r.grad = gr.clone(); o.grad = gr.clone()
        o_ref.step(); o_our.step()
    for r, o in zip(ref, ours):
        assert rel_err(o.detach().cpu(), r.detach().cpu()) < 1e-6
    # fused gradient zeroing
    o_z = P.FusedAdam(mk(ours), zero_grad_in_step=True)
    o_z.step()
    assert all(float(o.grad.abs().max()) == 0 for o in ours)
    # per-parameter step counts (torch keeps state['step'] per parameter): a tensor without a gradient in the first steps
    a_ref = [torch.randn(64, generator=g).to(DEV).requires_grad_(True) for _ in range(2)]
    a_our = [r.detach().clone().requires_grad_(True) for r in a_ref]
    t_ref, t_our = torch.optim.Adam(a_ref, lr=0.01), P.FusedAdam(a_our, lr=0.01)
    for it in range(4):
        for k, (r, o) in enumerate(zip(a_ref, a_our)):
            if k == 1 and it < 2:
                r.grad = None; o.grad = None                            # joins at step 3 with bias correction of ITS step 1
                continue
            gr = torch.randn(64, generator=g).to(DEV)
            r.grad = gr.clone(); o.grad = gr.clone()
        t_ref.step(); t_our.step()
    for r, o in zip(a_ref, a_our):
        assert rel_err(o.detach().cpu(), r.detach().cpu()) < 1e-6
    # checkpoints interchange with torch.optim.Adam
    t_our2 = P.FusedAdam([o.detach().clone().requires_grad_(True) for o in a_our], lr=0.01)
    t_our2.load_state_dict(t_ref.state_dict())                        # torch's checkpoint into ours
    t_our3 = P.FusedAdam([o.detach().clone().requires_grad_(True) for o in a_our], lr=0.01)
    t_our3.load_state_dict(t_our.state_dict())                        # and our own
    gr = torch.randn(64, generator=g).to(DEV)
    for opt_ in (t_ref, t_our2, t_our3):
        for p_ in opt_.param_groups[0]["params"]:
            p_.grad = gr.clone()
        opt_.step()
    for r, o2, o3 in zip(a_ref, t_our2.param_groups[0]["params"], t_our3.param_groups[0]["params"]):
        assert rel_err(o2.detach().cpu(), r.detach().cpu()) < 1e-6 and rel_err(o3.detach().cpu(), r.detach().cpu()) < 1e-6
    sd = t_our.state_dict()
    assert set(sd) == {"state", "param_groups"} and set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    with pytest.raises(ValueError):
        P.FusedAdam([torch.zeros(1, device=DEV, requires_grad=True) for _ in range(25)])   # the launch limit is reported at construction
    # reset_state() == a freshly built optimiser (the host code builds a new Adam per frame), also with the device step counter
    for dev_counter in (False, True):
        w0 = torch.randn(257, generator=g).to(DEV)
        pa, pb = w0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
        used, fresh = P.FusedAdam([pa], lr=0.02), P.FusedAdam([pb], lr=0.02)
        if dev_counter:
            used.enable_graph_step_counter(DEV); fresh.enable_graph_step_counter(DEV)
        for _ in range(3):
            pa.grad = torch.randn(257, generator=g).to(DEV); used.step()
        with torch.no_grad():
            pa.copy_(w0)
        used.reset_state()
        for _ in range(2):
            gr = torch.randn(257, generator=g).to(DEV)
            pa.grad = gr.clone(); pb.grad = gr.clone()
            used.step(); fresh.step()
        assert torch.equal(pa.detach(), pb.detach())


def test_tcgen05_forward_matches_cuda_core_forward(monkeypatch):
    """field_tc.cu (tangent contraction on tcgen05 tensor cores, TF32 operands) vs the all-fp32 CUDA-core kernel:
    identical raw outputs (value path is fp32 in both), Jacobian within the gradient tolerance."""
    P = pkg()
    g = load_golden("track_replica")
    meta, tabs, dec, beta = gpu_cases.cuda_field(g, 50, DEV)
    n = 128 * 37 + 5                                            # ragged last CTA
    x = torch.rand(n, 3, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3)) * 1.1 - 0.05   # some points clamp
    from ctypes import byref
    outs = []
    for entry in ("usl_field_fwd", "usl_field_fwd_tc"):
        raw = torch.zeros(n, 4, device=DEV); jac = torch.zeros(12, n, device=DEV)
        f = meta.pack(tabs[0], tabs[1], dec)
        pts = P.ops._points_from_x(x)
        P._lib.call(entry, byref(f), byref(pts), P._lib.ptr(raw), None, P._lib.ptr(jac), P._lib.stream())
        torch.cuda.synchronize()
        outs.append((raw.cpu(), jac.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) or (outs[0][0] - outs[1][0]).abs().max() < 1e-6
    assert rel_err(outs[1][1], outs[0][1]) < 1e-3
    assert (outs[1][1] - outs[0][1]).abs().max() < 2e-3 * outs[0][1].abs().max()


@pytest.mark.parametrize("mode", [0, 2])
def test_fused_composite_loss_forward_equals_the_two_calls(mode):
    """usl_composite_loss_fwd (one launch) vs usl_composite_fwd + usl_loss_fwd: identical per-ray outputs and masks, the
    same loss sums / counts; the tracker's mode (median mask) is refused."""
    from ctypes import byref
    P = pkg(); L = P._lib
    gen = torch.Generator(device=DEV).manual_seed(11 + mode)
    R, S = 1003, 40                                              # ragged last CTA
    raw = torch.rand(R, S, 4, device=DEV, generator=gen); raw[..., 3] = raw[..., 3] * 2 - 1
    z = (torch.rand(R, S, device=DEV, generator=gen) * 3).sort(dim=1).values.contiguous()
    beta = torch.tensor([8.0], device=DEV)
    valid = (torch.rand(R, device=DEV, generator=gen) > 0.1).to(torch.uint8)
    gt_d = torch.rand(R, device=DEV, generator=gen) * 3; gt_d[::17] = 0.0
    gt_c = torch.rand(R, 3, device=DEV, generator=gen)
    la = P.ops.make_loss_args(0.06, 5.0, 200.0, 10.0, 0.1, 5.0, mode)

    def outs():
        return [torch.full((R,), 7.0, device=DEV) for _ in range(3)] + [torch.full((R, 3), 7.0, device=DEV), torch.full((R,), 7.0, device=DEV),
                torch.zeros(16, device=DEV), torch.full((R,), 9, device=DEV, dtype=torch.uint8)]
    a = outs(); b = outs()
    st = L.stream()
    L.call("usl_composite_fwd", L.ptr(raw), L.ptr(z), L.ptr(beta), L.ptr(valid), R, S, L.ptr(a[0]), L.ptr(a[1]), L.ptr(a[2]), L.ptr(a[3]), L.ptr(a[4]), None, st)
    L.call("usl_loss_fwd", byref(la), L.ptr(raw), L.ptr(z), L.ptr(gt_d), L.ptr(gt_c), L.ptr(valid), L.ptr(a[1]), L.ptr(a[2]), L.ptr(a[3]), None, R, S,
           L.ptr(a[5]), L.ptr(a[6]), st)
    L.call("usl_composite_loss_fwd", byref(la), L.ptr(raw), L.ptr(z), L.ptr(beta), L.ptr(valid), R, S, L.ptr(gt_d), L.ptr(gt_c), L.ptr(b[0]), L.ptr(b[1]),
           L.ptr(b[2]), L.ptr(b[3]), L.ptr(b[4]), L.ptr(b[5]), L.ptr(b[6]), st)
    torch.cuda.synchronize()
    for i in (0, 1, 2, 3, 4, 6):
        assert torch.equal(a[i], b[i]), i
    assert float(a[5][5:11].sub(b[5][5:11]).abs().max()) == 0.0             # counts are exact
    assert rel_err(b[5], a[5]) < 1e-6                                         # sums: atomics reorder the additions
    la1 = P.ops.make_loss_args(0.06, 5.0, 200.0, 10.0, 0.1, 5.0, 1)
    with pytest.raises(RuntimeError, match="median"):
        L.call("usl_composite_loss_fwd", byref(la1), L.ptr(raw), L.ptr(z), L.ptr(beta), L.ptr(valid), R, S, L.ptr(gt_d), L.ptr(gt_c), L.ptr(b[0]),
               L.ptr(b[1]), L.ptr(b[2]), L.ptr(b[3]), L.ptr(b[4]), L.ptr(b[5]), L.ptr(b[6]), st)


def test_tracking_converges_on_held_out_frame():
    """End to end: fit the field on a keyframe window with MappingStep + FusedAdam, then TrackingStep + FusedAdam must pull a
    held-out frame's pose (perturbed by 2.7 cm / 1 deg) back to the analytic ground truth."""
    import json, subprocess, sys, os
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(repo, "tools", "track_convergence.py"), "replica_room0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["map_losses"][-1] < 0.05 * r["map_losses"][0]                 # mapping gradients train the field
    t0, r0 = r["track_err_cm_deg_every10"][0]
    t1, r1 = r["track_err_cm_deg_every10"][-1]
    assert t1 < 0.25 * t0 and t1 < 0.7 and r1 < 0.25 * r0, r             # pose gradients pull the camera back


def test_slam_loop_runs_and_stays_on_track():
    """BASELINE config 2 driver (uni-slam_b200/slam.py): short Tracker+Mapper loop on the synthetic Replica sequence."""
    import importlib
    slam = importlib.import_module("uni-slam_b200.slam")
    r = slam.run_slam(pkg().synthetic.REPLICA_ROOM0, n_frames=24, scale_hw=0.25)
    assert r.loss_last_map < 0.05 * r.loss_first_map
    assert r.ate_rmse < 0.03 and r.tracking_iters == 23 * 8 and r.mapping_iters == 10 + 5 * 15
    # the same loop with every iteration after the first of its shape replayed from a CUDA graph: same iteration counts,
    # same behaviour (RNG streams differ between captured and eager draws, so the trajectories are not bit-identical)
    g = slam.run_slam(pkg().synthetic.REPLICA_ROOM0, n_frames=24, scale_hw=0.25, graphs=True)
    assert g.loss_last_map < 0.05 * g.loss_first_map
    assert g.ate_rmse < 0.03 and g.tracking_iters == 23 * 8 and g.mapping_iters == 10 + 5 * 15
    assert abs(g.ate_rmse - r.ate_rmse) < 0.01


def test_empty_and_degenerate_inputs():
    """Edge cases of the seams: zero points / rays (tcnn pads an empty batch to nothing), an all-holes ray batch
    (n_valid == 0: the reference draws torch.rand((0,S)) and takes the no-depth branch for every ray), a batch in which
    no ray survives the mask (torch.mean over an empty selection: the loss is NaN in the reference),
    and argument errors raised as RuntimeError."""
    P = pkg()
    g = load_golden("map_replica_k7")
    meta, tabs, dec, beta = gpu_cases.cuda_field(g, 0, DEV)
    ns, ni, tr = int(g["n_stratified"]), int(g["n_importance"]), float(g["truncation"])
    S = ns + ni
    # -- zero-sized batches through every differentiable seam
    enc = P.Encoding(3, {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 16, "base_resolution": 16,
                         "per_level_scale": float(g["per_level_scale"][0])}, dtype=torch.float).to(DEV)
    x0 = torch.empty((0, 3), device=DEV, requires_grad=True)
    y0 = enc(x0)
    assert y0.shape == (0, 32)
    y0.sum().backward()
    assert enc.params.grad is not None and float(enc.params.grad.abs().sum()) == 0.0 and x0.grad.shape == (0, 3)
    assert P.ops.field_sdf_points(meta, torch.empty((0, 3), device=DEV), tabs[0], tabs[1], dec).shape == (0,)
    e3 = torch.empty((0, 3), device=DEV)
    ret = P.ops.render_rays(meta, e3, e3, torch.empty((0, S), device=DEV), beta, tabs[0], tabs[1], dec)
    assert ret[0].shape == (0,) and ret[3].shape == (0, 3) and ret[4].shape == (0, S)
    # -- all-holes batch through the fused renderer vs the oracle (no-depth branch for every ray)
    R = 37
    gen = torch.Generator().manual_seed(4)
    c2w = T(g["call0_c2ws"][1])
    dirs = T(g["call0_rays_d_cam"][1][:R])
    rays_d = torch.sum(dirs[:, None, :] * c2w[:3, :3], -1); rays_o = c2w[:3, 3].expand(R, 3).contiguous()
    t_uni, u_pdf = torch.rand((R, ns), generator=gen), torch.rand((R, ni), generator=gen)
    field = golden_field(g, 0, requires_grad=False)
    want = path_ref.render_batch_ray(field, rays_d, rays_o, torch.zeros(R), ns, ni, tr, torch.empty((0, S)), t_uni, u_pdf)
    zs = P.ops.ZSampler(ns, ni, tr, DEV)
    z = torch.zeros((R, S), device=DEV)
    gt0 = torch.zeros(R, device=DEV)
    zs.depth_guided(gt0, z, t_rand=torch.empty((0, S), device=DEV))
    zs.no_depth(meta.pack(tabs[0], tabs[1], dec), beta, rays_o.to(DEV), rays_d.to(DEV), gt0, z, u_pdf.to(DEV), t_rand_uni=t_uni.to(DEV))
    assert float((z.cpu() - want[5]).abs().max()) < 1e-4
    got = P.ops.render_rays(meta, rays_o.to(DEV), rays_d.to(DEV), z, beta, tabs[0], tabs[1], dec)
    assert max_rel(got[2].cpu(), want[2], 1e-3) < 1e-3 and max_rel(got[3].cpu(), want[3], 1e-3) < 1e-3
    # -- no ray survives the mask: NaN loss like torch.mean(empty), and nothing is scattered
    K = g["call0_c2ws"].shape[0]
    n = int(g["call0_n"])
    step = P.MappingStep(meta, tabs[0], tabs[1], dec, beta, n_stratified=ns, n_importance=ni, truncation=tr, max_rays=K * n, max_frames=K)
    far = torch.full_like(T(g["call0_depths"]), 1e4).to(DEV)                       # beyond the bbox exit: every ray is filtered
    batch = (T(g["call0_c2ws"]).to(DEV).contiguous(), far, T(g["call0_colors"]).to(DEV), T(g["call0_rays_d_cam"]).to(DEV),
             T(g["call0_indices"]).to(DEV), n, 0)
    loss = step.run([batch], torch.rand((K * n, S), device=DEV), torch.rand((K * n, ns), device=DEV), torch.rand((K * n, ni), device=DEV))
    assert int(step.valid[:K * n].sum()) == 0 and torch.isnan(loss).all()
    assert float(step.fs.g_sdf_table.abs().sum()) == 0.0 and float(step.fs.g_rgb_table.abs().sum()) == 0.0
    # -- argument errors
    with pytest.raises(RuntimeError):
        P.ops.ZSampler(120, 20, tr, DEV).depth_guided(torch.ones(4, device=DEV), torch.zeros((4, 140), device=DEV))   # S > 128
    with pytest.raises(RuntimeError):
        P.MappingStep(meta, tabs[0], tabs[1], dec, beta, n_stratified=ns, n_importance=ni, truncation=tr, max_rays=8).run(
            [batch], torch.rand((K * n, S), device=DEV))                              # more rays than the step was sized for


def test_median_select_and_keep_best():
    """torch.median (lower middle) of |gt - depth| over valid rays for ragged sizes, ties, NaN; and the tracker's
    on-device best-pose bookkeeping (Tracker.py:346-348)."""
    P = pkg()
    L = P._lib
    gen = torch.Generator().manual_seed(3)
    for n in (1, 2, 5, 255, 256, 2000, 4097):
        gt = torch.rand(n, generator=gen) * 3; dep = torch.rand(n, generator=gen) * 3
        if n >= 5:
            dep[1] = dep[0] = gt[0] - 0.25; gt[1] = gt[0]                            # ties
        valid = (torch.rand(n, generator=gen) > 0.2).to(torch.uint8)
        valid[0] = 1
        want = (gt - dep).abs()[valid.bool()].median()
        ws = torch.empty(n, device=DEV); med = torch.zeros(1, device=DEV)
        gd, dd, vd = gt.to(DEV), dep.to(DEV), valid.to(DEV)                         # keep the device copies alive across the call
        L.call("usl_depth_error_median", L.ptr(gd), L.ptr(dd), L.ptr(vd), n, L.ptr(ws), L.ptr(med), L.stream())
        assert float(med) == float(want), n                                          # selection: bit-exact
    dep[3] = float("nan"); valid[3] = 1
    gd, dd, vd = gt.to(DEV), dep.to(DEV), valid.to(DEV)
    L.call("usl_depth_error_median", L.ptr(gd), L.ptr(dd), L.ptr(vd), n, L.ptr(ws), L.ptr(med), L.stream())
    assert torch.isnan(med).all()                                                    # torch.median propagates NaN
    best_loss = torch.full((1,), float("inf"), device=DEV); best_pose = torch.zeros((1, 7), device=DEV)
    poses = torch.randn(4, 1, 7, device=DEV)
    for loss, keep in ((3.0, True), (5.0, False), (float("nan"), False), (1.0, True)):
        i = [3.0, 5.0, float("nan"), 1.0].index(loss) if loss == loss else 2
        lt = torch.tensor([loss], device=DEV)
        L.call("usl_track_keep_best", L.ptr(lt), L.ptr(poses[i]), L.ptr(best_loss), L.ptr(best_pose), L.stream())
        assert torch.equal(best_pose, poses[i]) == keep
    assert float(best_loss) == 1.0


def test_peer_collectives_single_rank_and_sharded_adam():
    """csrc/collective.cu on one GPU (world = 1: the peer is this rank itself): the loss-sum exchange and the all-reduce are
    identities, and usl_allreduce_adam_step reproduces torch.optim.Adam (two learning-rate ranges, several steps, device-side
    step counter).  The N > 1 behaviour is checked inside bench.py --gpus N ("allreduce_check")."""
    import importlib
    P = pkg()
    par = importlib.import_module("uni-slam_b200.parallel")
    pg = par.PeerGroup.single(DEV)
    g = torch.Generator(device=DEV).manual_seed(3)
    acc = torch.rand(16, device=DEV, generator=g)
    want = acc.clone()
    for _ in range(3):                                   # epochs advance; double-buffered slots
        pg.exchange_sums(acc)
    torch.cuda.synchronize()
    assert torch.equal(acc, want)
    n = 4096 * 5 + 64
    grads = pg.alloc(n); params = pg.alloc(n)
    grads.copy_(torch.randn(n, device=DEV, generator=g))
    before = grads.clone()
    pg.allreduce(grads, n)
    pg.barrier()
    torch.cuda.synchronize()
    assert torch.equal(grads, before)
    params.copy_(torch.randn(n, device=DEV, generator=g))
    split = 4096 * 3
    ref = [params[:split].clone().requires_grad_(True), params[split:n - 64].clone().requires_grad_(True)]
    opt = torch.optim.Adam([{"params": [ref[0]], "lr": 0.05}, {"params": [ref[1]], "lr": 1e-3}])
    fsa = par.FusedShardedAdam(pg, params, grads, n, [(0, split, 0.05), (split, n - 64, 1e-3)])
    tail = params[n - 64:].clone()
    for it in range(4):
        gr = torch.randn(n, device=DEV, generator=g) * (0.1 + it)
        grads.copy_(gr)
        ref[0].grad = gr[:split].clone(); ref[1].grad = gr[split:n - 64].clone()
        opt.step()
        fsa.step()
    torch.cuda.synchronize()
    assert rel_err(params[:split].cpu(), ref[0].detach().cpu()) < 1e-6
    assert rel_err(params[split:n - 64].cpu(), ref[1].detach().cpu()) < 1e-6
    assert torch.equal(params[n - 64:], tail)            # floats outside every learning-rate range are left untouched


def test_keyframe_store_on_device_matches_reference():
    """f2 on the GPU: the device-resident KeyframeStore (usl_keyframe_insert) holds the tensors the unmodified reference stacks
    (fixture map_replica_kfstore), a MappingStep fed from its views reproduces the reference's loss, and the co-visibility
    measure of keyframe_selection_LC (usl_keyframe_covisibility) matches the reference's percent_inside."""
    P = pkg()
    g = load_golden("map_replica_kfstore")
    H, W = int(g["meta_H_W_fx_fy_cx_cy"][0]), int(g["meta_H_W_fx_fy_cx_cy"][1])
    col, dep, gtc = T(g["frames_color"]).to(DEV), T(g["frames_depth"]).to(DEV), T(g["frames_gt_c2w"]).to(DEV)
    dirs = T(g["dirs_cam"]).to(DEV)
    n_kf = g["kf_indices"].shape[0]
    store = P.KeyframeStore(capacity=8, H=H, W=W, device=DEV)
    for k in range(n_kf):
        store.append(int(g["kf_frame_idx"][k]), col[k], dep[k], dirs, T(g["kf_est_c2w"][k]).to(DEV), gtc[k], indices=T(g["kf_indices"][k]).to(DEV))
    store.stage_current(col[n_kf], dep[n_kf], dirs, T(g["cur_c2w"]).to(DEV), gtc[n_kf], indices=T(g["cur_randperm"])[:store.P].to(DEV))
    c2ws, depths, colors, rays_d = store.window()
    for got, key in ((c2ws, "call0_c2ws"), (depths, "call0_depths"), (colors, "call0_colors"), (rays_d, "call0_rays_d_cam")):
        assert torch.equal(got.cpu(), T(g[key])), key
    assert torch.equal(store.pixel_idx[0].cpu(), T(g["kf_indices"][0]))
    # the mapping iteration straight from the store's views
    meta, tabs, dec, beta = gpu_cases.cuda_field(g, 0, DEV)
    ns, ni = int(g["n_stratified"]), int(g["n_importance"]); S = ns + ni
    n = int(g["call0_n"])
    batches = store.mapping_batches(T(g["call0_indices"]).to(DEV), n)
    R = depths.shape[0] * n
    ro = T(g["call0_out_rays_o"]); rd = T(g["call0_out_rays_d"]); dd = T(g["call0_out_depth"])
    inside = path_ref.bbox_exit(ro, rd, T(g["bound"])) >= dd
    t_rand, t_uni, u_pdf, has_holes = gpu_cases._slot_draws(g, inside, dd, S, ns, ni, DEV)
    step = P.MappingStep(meta, tabs[0], tabs[1], dec, beta, n_stratified=ns, n_importance=ni, truncation=float(g["truncation"]), max_rays=R, max_frames=8)
    loss = step.run(batches, t_rand, t_uni, u_pdf, has_holes=has_holes)
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-4
    # co-visibility
    c = load_golden("kf_covis_replica")
    Hc, Wc, fx, fy, cx, cy = [float(v) for v in c["meta_H_W_fx_fy_cx_cy"]]
    kc = T(c["keyframe_c2ws"]).to(DEV)
    st2 = P.KeyframeStore(capacity=12, H=int(Hc), W=int(Wc), device=DEV)
    for k in range(kc.shape[0]):
        st2.est_c2w[k] = kc[k]; st2.frame_idx.append(4 * k)
    pi = st2.covisibility(T(c["sample_out_rays_o"]).to(DEV), T(c["sample_out_rays_d"]).to(DEV), T(c["sample_out_depth"]).to(DEV),
                          int(Hc), int(Wc), fx, fy, cx, cy, int(c["num_samples"]), float(c["edge"]))
    ref = T(c["percent_inside"])
    n_pts = int((T(c["sample_out_depth"]) > 0).sum()) * int(c["num_samples"])
    assert (pi.cpu() - ref).abs().max() <= 1.5 / n_pts                  # at most one borderline point (fp32 inverse vs closed form)
