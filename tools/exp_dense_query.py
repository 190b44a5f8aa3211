import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
P = importlib.import_module("uni-slam_b200")
wlmod = importlib.import_module("uni-slam_b200.workload")
dev = "cuda:0"
cfg = P.synthetic.REPLICA_ROOM0
bound = P.synthetic.load_bound(cfg.bound_yaml)
pls = float(np.exp2(np.log2(816 / 16) / 15))
meta, tabs, dec, beta = wlmod.init_field_tensors(cfg, bound, pls, dev, seed=0)
tabs = [torch.randn_like(t) * 0.05 for t in tabs]
axes = []
for a in range(3):
    lo, hi = cfg.bound_yaml[a]
    axes.append(torch.from_numpy(np.linspace(lo - 0.05, hi + 0.05, int(round((hi - lo + 0.1) / 0.01)))).float().to(dev))
q = P.DenseSdfQuery(meta, tabs[0], tabs[1], dec, axes)
out = torch.empty(q.slab_points(0, q.ny), device=dev)
for _ in range(2): q.run(0, q.ny, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): q.run(0, q.ny, out)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("USL_LIB_PATH", "default").split("/")[-1], "dense query ms", e0.elapsed_time(e1) / 3, "checksum", float(out[::1000].double().sum()))
