"""CPU tests: the C-ABI library loads and exports every symbol the header declares; host-side logic
(level tables, module construction / pickling, integration wiring) works without a GPU."""
import copy
import importlib
import os
import pickle
import re
import sys

import numpy as np
import pytest
import torch

from oracle import grid_ref
from helpers import pkg

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_abi_exports_match_header():
    P = pkg()
    lib = P._lib.load()
    hdr = open(os.path.join(REPO, "include", "unislam_b200.h")).read()
    declared = set(re.findall(r"USL_API\s+(?:const\s+char\s*\*|int)\s*(usl_\w+)\s*\(", hdr))
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/unislam_b200.h but not exported"
    assert declared == set(P._lib.EXPORTS), declared ^ set(P._lib.EXPORTS)
    assert lib.usl_version() >= 100


@pytest.mark.parametrize("log2T,res", [(16, 816), (19, 816), (16, 456)])
def test_level_table_matches_oracle(log2T, res):
    P = pkg()
    pls = grid_ref.per_level_scale_from_resolution(res)
    g = P._lib.build_grid(16, log2T, 16, pls)
    spec = grid_ref.make_grid_spec(log2T, pls)
    assert g.total_entries == spec.total_entries
    for a, b in zip(list(g.levels)[:16], spec.levels):
        assert (a.scale, a.res, a.size, a.offset, bool(a.hashed)) == (np.float32(b.scale), b.res, b.size, b.offset, b.hashed)


def test_grid_build_rejects_bad_arguments():
    P = pkg()
    with pytest.raises(RuntimeError):
        P._lib.build_grid(17, 16, 16, 1.3)
    with pytest.raises(RuntimeError):
        P._lib.build_grid(16, 40, 16, 1.3)


def test_modules_construct_pickle_and_refuse_cpu():
    P = pkg()
    enc = P.Encoding(3, {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 16,
                         "base_resolution": 16, "per_level_scale": 1.2996847159335432}, dtype=torch.float)
    assert enc.params.numel() == 868400 * 2 and enc.n_output_dims == 32 and enc.params.is_leaf
    e2 = pickle.loads(pickle.dumps(enc)); e3 = copy.deepcopy(enc)
    assert torch.equal(e2.params, enc.params) and torch.equal(e3.params, enc.params)
    assert e2.grid.total_entries == enc.grid.total_entries == e3.grid.total_entries
    with pytest.raises(RuntimeError):
        enc(torch.rand(4, 3))
    net = P.Network(32, 3, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "Sigmoid", "n_neurons": 16, "n_hidden_layers": 1})
    assert net.params.numel() == 768
    with pytest.raises(RuntimeError):
        net(torch.rand(4, 32))
    with pytest.raises(RuntimeError):
        P.Network(32, 3, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "Sigmoid", "n_neurons": 64, "n_hidden_layers": 1})
    for tc in (True, False):
        cfg = {"grid_mode": "hash_grid", "grid": {"tcnn_network": tc}}
        dec = P.Decoders(cfg, c_dim=32, truncation=0.06, learnable_beta=True)
        keys = set(dec.state_dict().keys())
        want = {"beta", "sdf_decoder.params", "color_decoder.params"} if tc else \
            {"beta"} | {f"{a}.{i}.{w}" for a in ("linears", "c_linears") for i in (0, 1) for w in ("weight", "bias")} | \
            {f"{a}.{w}" for a in ("output_linear", "c_output_linear") for w in ("weight", "bias")}
        assert keys == want


def test_missing_library_fails_loudly(monkeypatch):
    P = pkg()
    monkeypatch.setattr(P._lib, "_lib", None)
    monkeypatch.setattr(P._lib, "LIB_PATH", "/nonexistent/libunislam_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P._lib.load()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_reference_constructs_our_modules_through_compat():
    """The UNMODIFIED reference Decoders / get_encoder arithmetic build B200 modules when compat/ is first on sys.path."""
    code = r'''
import sys, importlib
sys.path[:0] = ["%s/uni-slam_b200/compat", "%s/oracle/shims", "%s", "%s"]
import tinycudann as tcnn
P = importlib.import_module("uni-slam_b200")
assert tcnn.Encoding is P.Encoding and tcnn.Network is P.Network
from src.networks.decoders import Decoders
cfg = {"grid_mode": "hash_grid", "grid": {"tcnn_network": True}}
d = Decoders(cfg, c_dim=32, truncation=0.06, learnable_beta=True)
assert isinstance(d.sdf_decoder, P.Network) and sorted(d.state_dict()) == ["beta", "color_decoder.params", "sdf_decoder.params"]
import numpy as np, torch
pls = np.exp2(np.log2(816 / 16) / 15)
e = tcnn.Encoding(n_input_dims=3, encoding_config={"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2,
    "log2_hashmap_size": 19, "base_resolution": 16, "per_level_scale": pls}, dtype=torch.float)
assert e.n_output_dims == 32 and e.params.numel() == 5588448 * 2
print("ok")
''' % (REPO, REPO, REF, REPO)
    import subprocess
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_keyframe_store_matches_reference_bookkeeping():
    """f2: the device-resident KeyframeStore must hand MappingStep the very tensors the unmodified reference's
    optimize_mapping stacks from its keyframe_dict + the current frame's randperm subset (Mapper.py:315-351;
    fixture recorded from the reference by oracle/gen_golden.py), as views, and sampling from them must reproduce
    get_samples_all bit for bit."""
    from oracle import path_ref
    from helpers import load_golden
    KeyframeStore = importlib.import_module("uni-slam_b200.keyframes").KeyframeStore
    g = load_golden("map_replica_kfstore")
    T = torch.from_numpy
    H, W = int(g["meta_H_W_fx_fy_cx_cy"][0]), int(g["meta_H_W_fx_fy_cx_cy"][1])
    col, dep, gtc = T(g["frames_color"]), T(g["frames_depth"]), T(g["frames_gt_c2w"])
    dirs = T(g["dirs_cam"])
    n_kf = g["kf_indices"].shape[0]
    store = KeyframeStore(capacity=8, H=H, W=W, device="cpu")
    assert store.P == g["kf_indices"].shape[1] == int(H * W * 0.1)
    for k in range(n_kf):
        store.append(int(g["kf_frame_idx"][k]), col[k], dep[k], dirs, T(g["kf_est_c2w"][k]), gtc[k], indices=T(g["kf_indices"][k]))
    store.stage_current(col[n_kf], dep[n_kf], dirs, T(g["cur_c2w"]), gtc[n_kf], indices=T(g["cur_randperm"])[:store.P])
    c2ws, depths, colors, rays_d = store.window()
    for got, key in ((c2ws, "call0_c2ws"), (depths, "call0_depths"), (colors, "call0_colors"), (rays_d, "call0_rays_d_cam")):
        assert torch.equal(got, T(g[key])), key
        assert got.untyped_storage().data_ptr() in {t.untyped_storage().data_ptr() for t in (store.est_c2w, store.depth, store.color, store.rays_d)}   # a view, not a copy
    (cw, dp, cl, rd, idx, n, base), = store.mapping_batches(T(g["call0_indices"]), int(g["call0_n"]))
    out = path_ref.sample_mapping_rays(cw, dp, cl, rd, idx)
    for nm, t in zip(("rays_o", "rays_d", "depth", "color"), out):
        assert torch.equal(t, T(g["call0_out_" + nm])), nm
    # wire format + non-contiguous selection (loop-closure style) + pose write-back
    d = store.as_dicts()
    assert [x["idx"] for x in d] == list(g["kf_frame_idx"]) and set(d[0]) == {"gt_c2w", "idx", "color", "depth", "est_c2w", "rays_d"}
    sel = store.window([0, 2])
    assert torch.equal(sel[1], torch.stack([store.depth[0], store.depth[2], store.depth[n_kf]]))
    new = torch.randn(n_kf, 4, 4)
    cur = store.write_back_poses(new)
    assert torch.equal(store.est_c2w[1], new[0]) and torch.equal(store.est_c2w[2], new[1]) and torch.equal(cur, new[-1])
    assert torch.equal(store.est_c2w[0], T(g["kf_est_c2w"][0]))                    # the oldest frame stays fixed
    with pytest.raises(RuntimeError):
        for k in range(20):
            store.promote_staged(100 + k)


def test_workload_draw_buffers_and_mask_modes():
    """Host logic of the fused drivers: the two flat draw buffers expose the five slot-indexed tensors as views in draw()
    order (one index fill + one uniform fill reach all of them), and the loss mask modes map onto the C-ABI enum."""
    wl = importlib.import_module("uni-slam_b200.workload")
    syn = importlib.import_module("uni-slam_b200.synthetic")
    steps = importlib.import_module("uni-slam_b200.steps")
    w = wl.build_mapping_workload(syn.REPLICA_ROOM0, "cpu", n_keyframes=21, scale_hw=0.05)
    flat_idx, flat_u, (idx_main, idx_recent, t_rand, t_uni, u_pdf) = w.alloc_draws()
    assert idx_main.shape == (w.K * w.n_main,) and idx_recent.shape == (10 * w.n_recent,)
    assert t_rand.shape == (w.n_rays, w.S) and t_uni.shape == (w.n_rays, 32) and u_pdf.shape == (w.n_rays, 8)
    ref = w.draw(torch.Generator().manual_seed(0))
    assert [tuple(t.shape) for t in ref] == [tuple(t.shape) for t in (idx_main, idx_recent, t_rand, t_uni, u_pdf)]
    flat_idx.fill_(7); flat_u.fill_(0.5)
    assert all(bool((t == 7).all()) for t in (idx_main, idx_recent)) and all(bool((t == 0.5).all()) for t in (t_rand, t_uni, u_pdf))
    assert flat_idx.numel() == idx_main.numel() + idx_recent.numel()
    assert flat_u.numel() == t_rand.numel() + t_uni.numel() + u_pdf.numel()
    w1 = wl.build_mapping_workload(syn.REPLICA_ROOM0, "cpu", n_keyframes=3, scale_hw=0.05)   # <= 20 keyframes: no recent-frame batch
    assert w1.alloc_draws()[2][1] is None and w1.n_recent == 0
    assert steps._mask_mode("original", 0) == 0 and steps._mask_mode("original", 1) == 1 and steps._mask_mode("no_mask", 0) == 2
    with pytest.raises(ValueError):
        steps._mask_mode("alpha", 0)


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """Every struct the Python layer passes by pointer must have the C compiler's size and field offsets: compile a probe
    against include/unislam_b200.h with gcc and compare with ctypes."""
    import ctypes
    import subprocess
    L = importlib.import_module("uni-slam_b200._lib")
    pairs = [("usl_level_t", L.Level), ("usl_grid_t", L.Grid), ("usl_mlp_t", L.Mlp), ("usl_field_t", L.Field), ("usl_bound_t", L.Bound),
             ("usl_points_t", L.Points), ("usl_zsample_args_t", L.ZSampleArgs), ("usl_ray_batch_t", L.RayBatch), ("usl_ray_setup_t", L.RaySetup),
             ("usl_loss_args_t", L.LossArgs), ("usl_adam_group_t", L.AdamGroup), ("usl_peers_t", L.Peers), ("usl_adam_range_t", L.AdamRange),
             ("usl_mc_args_t", L.McArgs), ("usl_cull_frames_args_t", L.CullFramesArgs)]
    last = {"usl_ray_setup_t": "ray_offset", "usl_peers_t": "max_ctas_per_sm", "usl_mc_args_t": "faces", "usl_adam_group_t": "step",
            "usl_points_t": "n", "usl_field_t": "bound_hi", "usl_cull_frames_args_t": "seen"}
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(REPO, "include", "unislam_b200.h")}"', "int main(void) {"]
    for cname, _ in pairs:
        src.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
    for cname, fld in last.items():
        src.append(f'  printf("{cname}.{fld} %zu\\n", offsetof({cname}, {fld}));')
    src.append("  return 0; }")
    c = tmp_path / "probe.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(c)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines())
    for cname, cls in pairs:
        assert int(out[cname]) == ctypes.sizeof(cls), cname
    byname = dict(pairs)
    for cname, fld in last.items():
        assert int(out[f"{cname}.{fld}"]) == getattr(byname[cname], fld).offset, (cname, fld)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_aligned_ate_matches_the_reference_evaluation():
    """slam.ate_rmse_aligned against the unmodified align() / rmse of src/tools/eval_ate.py on a noisy, rotated trajectory."""
    import subprocess
    import json
    code = r"""
import sys, os, json
sys.path[:0] = [os.path.join(%r, "oracle", "shims"), %r, %r]
os.chdir(%r)
import numpy, torch
import src.tools.eval_ate as EA
rng = numpy.random.default_rng(4)
n = 57
gt = numpy.cumsum(rng.normal(0, 0.05, (n, 3)), axis=0)
a = 0.3
R = numpy.array([[numpy.cos(a), -numpy.sin(a), 0], [numpy.sin(a), numpy.cos(a), 0], [0, 0, 1.0]])
est = gt @ R.T + numpy.array([0.4, -0.2, 0.1]) + rng.normal(0, 0.01, (n, 3))
rot, trans, te = EA.align(numpy.matrix(est.T), numpy.matrix(gt.T))
rmse = float(numpy.sqrt(numpy.dot(te, te) / len(te)))
print("RES " + json.dumps({"gt": gt.tolist(), "est": est.tolist(), "rot": numpy.asarray(rot).tolist(), "trans": numpy.asarray(trans).ravel().tolist(), "rmse": rmse, "te": te.tolist()}))
""" % (REPO, REF, REPO, REF)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    res = json.loads([ln for ln in out.splitlines() if ln.startswith("RES ")][0][4:])
    slam = importlib.import_module("uni-slam_b200.slam")
    def c2w(xyz):
        m = torch.eye(4, dtype=torch.float64).repeat(len(xyz), 1, 1); m[:, :3, 3] = torch.tensor(xyz, dtype=torch.float64); return m
    rmse, rot, trans, te = slam.ate_rmse_aligned(c2w(res["est"]), c2w(res["gt"]))
    assert abs(rmse - res["rmse"]) < 1e-12 and np.allclose(rot, res["rot"], atol=1e-12) and np.allclose(trans, res["trans"], atol=1e-12)
    assert np.allclose(te, res["te"], atol=1e-12) and 0.005 < rmse < 0.03            # the 1 cm noise, not the 0.4 m offset


def test_host_side_entry_points_and_argument_errors():
    """Entry points that compute on the host, and the argument checks of launch wrappers that return BEFORE any CUDA call
    (status 1 + usl_last_error, status 0 for empty work) -- callable without a GPU."""
    from ctypes import byref, c_int64
    L = importlib.import_module("uni-slam_b200._lib")
    lib = L.load()
    # slices of the sharded Adam: multiples of 4 floats that cover the buffer
    for world in (1, 2, 3, 8):
        for n in (4, 1024, 12913796 // 4 * 4):
            s = c_int64(0)
            assert lib.usl_allreduce_adam_slice_floats(world, n, byref(s)) == 0
            assert s.value % 4 == 0 and s.value * world >= n and (s.value - 4) * world < n
    assert lib.usl_allreduce_adam_slice_floats(9, 1024, byref(s)) == 1 and b"bad arguments" in lib.usl_last_error()
    assert lib.usl_allreduce_adam_slice_floats(2, 1023, byref(s)) == 1
    nb = c_int64(0)
    assert lib.usl_scan_u8_blocks(0, byref(nb)) == 0 and nb.value == 0
    assert lib.usl_scan_u8_blocks(1, byref(nb)) == 0 and nb.value == 1
    assert lib.usl_scan_u8_blocks(-1, byref(nb)) == 1
    # mesh culling / metrics: empty work is a no-op, missing buffers are refused with a message
    a = L.CullFramesArgs()
    assert lib.usl_mesh_cull_frames(None, None) == 1
    assert lib.usl_mesh_cull_frames(byref(a), None) == 0                       # V = 0, K = 0
    a.V, a.K, a.H, a.W = 10, 2, 4, 4
    assert lib.usl_mesh_cull_frames(byref(a), None) == 1 and b"usl_mesh_cull_frames" in lib.usl_last_error()
    assert lib.usl_mesh_cull_hull(None, 0, None, 3, None, None) == 0
    assert lib.usl_mesh_cull_hull(None, 5, None, 3, None, None) == 1 and b"usl_mesh_cull_hull" in lib.usl_last_error()
    assert lib.usl_mesh_face_keep(None, 0, None, 0, 0, None, None, None) == 0
    assert lib.usl_mesh_face_keep(None, 7, None, 3, 0, None, None, None) == 1
    assert lib.usl_mesh_compact(None, None, 0, None, 0, None, None, None, None, None, None, None, None) == 0
    assert lib.usl_mesh_compact(None, None, 4, None, 0, None, None, None, None, None, None, None, None) == 1
    assert lib.usl_render_metrics(None, None, None, None, 0, None, None) == 0
    assert lib.usl_render_metrics(None, None, None, None, 9, None, None) == 1 and b"usl_render_metrics" in lib.usl_last_error()
    with pytest.raises(RuntimeError, match="usl_render_metrics"):
        L.call("usl_render_metrics", None, None, None, None, 9, None, None)
