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
