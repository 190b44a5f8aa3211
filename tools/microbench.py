"""GPU microbenchmarks that ground the roofline discussion.  Writes one JSON document:

  ceilings.l2_read_gbs            coalesced reads of an L2-resident 32 MB buffer (ld.global.cg, 16 B per lane)
  ceilings.hbm_read_gbs           the same kernel over a 1 GiB buffer
  ceilings.l1_sector_lookups_per_s   scattered 8-byte loads into an L2-resident table, every lane its own sector:
                                     the L1TEX tag stage looks up one 32-byte sector per clock per SM
  ceilings.atomic_sectors_per_s      scattered 8-byte vector atomics (red.global.add.v2.f32), every lane its own sector
  ceilings.atomic_sectors_per_s_pairs / _quads / _warp: lanes sharing a 16 B slot / a 32 B sector / 256 contiguous bytes

plus the raw sweeps (thread counts swept until the rate saturates) and the stand-alone encode kernels on ray-coherent vs
random points.  bench.py reads profiles/rNN_microbench.json for its L2 / L1TEX / atomic fractions.
Run on the GPU box:  python tools/microbench.py > gpurun_out/r02_microbench.json"""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
P = importlib.import_module("uni-slam_b200")
L = P._lib
dev = "cuda:0"


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {"gpu": torch.cuda.get_device_name(0), "sm_count": torch.cuda.get_device_properties(0).multi_processor_count,
       "l2_bytes": torch.cuda.get_device_properties(0).L2_cache_size}
out = torch.zeros(1 << 22, device=dev)

# ---- streaming reads: L2-resident and HBM ----
stream = {}
for mb, reps in ((16, 40), (32, 40), (64, 20), (1024, 2)):
    buf = torch.zeros(mb * 1024 * 1024 // 4, device=dev)
    ms = timeit(lambda: L.call("usl_bench_stream_read", L.ptr(buf), buf.numel(), reps, L.ptr(out), L.stream()), n=5)
    stream[f"{mb}MB"] = {"ms": ms, "GBs": buf.numel() * 4 * reps / ms / 1e6, "repeats": reps}
    del buf
res["stream_read"] = stream

# ---- scattered loads / atomics, swept over the number of threads until the rate stops growing ----
sweep = {}
for mb in (6, 42):
    entries = mb * 1024 * 1024 // 8
    table = torch.zeros(entries * 2, device=dev)
    for log2n in (18, 19, 20, 21, 22):
        nthr, per = 1 << log2n, 32
        ms = timeit(lambda: L.call("usl_bench_gather", L.ptr(table), entries, nthr, per, L.ptr(out), L.stream()))
        sweep[f"gather_{mb}MB_2^{log2n}"] = {"ms": ms, "G_per_s": nthr * per / ms / 1e6}
        ms = timeit(lambda: L.call("usl_bench_scatter", L.ptr(table), entries, nthr, per, 0, L.stream()))
        sweep[f"atomic_{mb}MB_2^{log2n}"] = {"ms": ms, "G_per_s": nthr * per / ms / 1e6}
    nthr, per = 1 << 21, 32
    for mode, nm, lanes_per_unit in ((1, "pairs16B", 2), (3, "quads32B", 4), (4, "warp256B", 4), (2, "float4", 1), (5, "same32", 32), (6, "same4", 4), (7, "same2", 2)):
        ms = timeit(lambda: L.call("usl_bench_scatter", L.ptr(table), entries, nthr, per, mode, L.stream()))
        sweep[f"atomic_{mb}MB_{nm}"] = {"ms": ms, "G_lane_atomics_per_s": nthr * per / ms / 1e6, "G_sectors_per_s": nthr * per / lanes_per_unit / ms / 1e6}
    del table
res["sweep"] = sweep
best = lambda pre: max(v["G_per_s"] for k, v in sweep.items() if k.startswith(pre))
res["ceilings"] = {
    "l2_read_gbs": max(stream["16MB"]["GBs"], stream["32MB"]["GBs"]),
    "hbm_read_gbs": stream["1024MB"]["GBs"],
    "l1_sector_lookups_per_s": best("gather_42MB_2^") * 1e9,
    "atomic_sectors_per_s": best("atomic_42MB_2^") * 1e9,
    "atomic_sectors_per_s_pairs": sweep["atomic_42MB_pairs16B"]["G_sectors_per_s"] * 1e9,
    "atomic_sectors_per_s_quads": sweep["atomic_42MB_quads32B"]["G_sectors_per_s"] * 1e9,
    "atomic_sectors_per_s_warp": sweep["atomic_42MB_warp256B"]["G_sectors_per_s"] * 1e9,
}

# ---- stand-alone encode kernels, Replica grids, 240k points ----
syn = P.synthetic
cfg = syn.REPLICA_ROOM0
bound = syn.load_bound(cfg.bound_yaml)
import numpy as np
from ctypes import byref
pls = float(np.exp2(np.log2(816 / 16) / 15))
n = 5982 * 40
xr = torch.rand(n, 3, device=dev)
o = torch.tensor([3.0, 1.2, -0.2], device=dev)
d = torch.randn(5982, 3, device=dev); d = d / d.norm(dim=-1, keepdim=True)
t = torch.linspace(0.05, 2.5, 40, device=dev)
pts = o + d[:, None, :] * t[None, :, None]
xc = ((pts - bound[:, 0].to(dev)) / (bound[:, 1] - bound[:, 0]).to(dev)).clamp(0, 1).reshape(-1, 3).contiguous()
dy = torch.randn(n, 32, device=dev)
enc = {}
for name, log2T in (("sdf16", 16), ("rgb19", 19)):
    g = L.build_grid(16, log2T, 16, pls)
    params = torch.randn(g.total_entries * 2, device=dev) * 0.05
    grad = torch.zeros_like(params)
    y = torch.empty(n, 32, device=dev)
    for pname, x in (("random", xr), ("rays", xc)):
        ms_f = timeit(lambda: L.call("usl_grid_encode_fwd", byref(g), L.ptr(params), L.ptr(x), n, L.ptr(y), L.stream()))
        ms_b = timeit(lambda: L.call("usl_grid_encode_bwd_params", byref(g), L.ptr(x), L.ptr(dy), n, L.ptr(grad), L.stream()))
        enc[f"{name}_{pname}"] = {"fwd_ms": ms_f, "bwd_params_ms": ms_b, "fwd_GBs_alg": n * 1164 / ms_f / 1e6, "bwd_GBs_alg": n * 1164 / ms_b / 1e6}
res["encode"] = enc
print(json.dumps(res, indent=1))
