import importlib, json, os, sys
sys.path.insert(0, "/root/repo")
import torch
P = importlib.import_module("uni-slam_b200"); L = P._lib
dev="cuda:0"
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for kb in (32, 74, 175, 512, 4096, 43008):
    entries = kb * 1024 // 8
    table = torch.zeros(entries * 2, device=dev)
    nthr, per = 1 << 21, 32
    for mode, nm in ((0, "random"), (2, "float4")):
        ms = timeit(lambda: L.call("usl_bench_scatter", L.ptr(table), entries, nthr, per, mode, L.stream()))
        print(f"table {kb:6d} KB {nm:8s}: {nthr*per/ms/1e6:7.1f} G lane-atomics/s")
