"""GPU microbenchmarks that ground the roofline discussion: L2-resident random 8-byte
gather and vector-atomic scatter ceilings, and the stand-alone encode kernels on ray-coherent vs random points."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
P = importlib.import_module("uni-slam_b200")
L = P._lib
dev = "cuda:0"

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {}
out = torch.zeros(1 << 22, device=dev)
for mb in (6, 42):
    entries = mb * 1024 * 1024 // 8
    table = torch.zeros(entries * 2, device=dev)
    nthr, per = 1 << 21, 32
    ms = timeit(lambda: L.call("usl_bench_gather", L.ptr(table), entries, nthr, per, L.ptr(out), L.stream()))
    res[f"gather_{mb}MB"] = {"ms": ms, "G_loads_per_s": nthr * per / ms / 1e6, "GBs_8B": nthr * per * 8 / ms / 1e6}
    for mode, nm in ((0, "random"), (1, "pairs16B"), (3, "quads32B"), (4, "warp256B"), (2, "float4"), (5, "same32"), (6, "same4"), (7, "same2")):
        ms = timeit(lambda: L.call("usl_bench_scatter", L.ptr(table), entries, nthr, per, mode, L.stream()))
        res[f"scatter_{mb}MB_{nm}"] = {"ms": ms, "G_lane_atomics_per_s": nthr * per / ms / 1e6}

# stand-alone encode kernels, Replica grids, 240k points
wl = importlib.import_module("uni-slam_b200.workload")
syn = P.synthetic
cfg = syn.REPLICA_ROOM0
bound = syn.load_bound(cfg.bound_yaml)
import numpy as np
pls = float(np.exp2(np.log2(816 / 16) / 15))
n = 5982 * 40
xr = torch.rand(n, 3, device=dev)
# ray-coherent points: rays from a room centre, 40 samples along each
o = torch.tensor([3.0, 1.2, -0.2], device=dev)
d = torch.randn(5982, 3, device=dev); d = d / d.norm(dim=-1, keepdim=True)
t = torch.linspace(0.05, 2.5, 40, device=dev)
pts = o + d[:, None, :] * t[None, :, None]
xc = ((pts - bound[:, 0].to(dev)) / (bound[:, 1] - bound[:, 0]).to(dev)).clamp(0, 1).reshape(-1, 3).contiguous()
dy = torch.randn(n, 32, device=dev)
for name, log2T in (("sdf16", 16), ("rgb19", 19)):
    g = L.build_grid(16, log2T, 16, pls)
    params = torch.randn(g.total_entries * 2, device=dev) * 0.05
    grad = torch.zeros_like(params)
    y = torch.empty(n, 32, device=dev)
    from ctypes import byref
    for pname, x in (("random", xr), ("rays", xc)):
        ms_f = timeit(lambda: L.call("usl_grid_encode_fwd", byref(g), L.ptr(params), L.ptr(x), n, L.ptr(y), L.stream()))
        ms_b = timeit(lambda: L.call("usl_grid_encode_bwd_params", byref(g), L.ptr(x), L.ptr(dy), n, L.ptr(grad), L.stream()))
        res[f"encode_{name}_{pname}"] = {"fwd_ms": ms_f, "bwd_params_ms": ms_b, "fwd_GBs_alg": n * 1164 / ms_f / 1e6, "bwd_GBs_alg": n * 1164 / ms_b / 1e6}
print(json.dumps(res, indent=1))
