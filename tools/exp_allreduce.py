"""torchrun experiment: the gradient exchange alone -- NCCL all_reduce vs the hand-written peer-memory kernels
(usl_allreduce_sum, usl_allreduce_adam_step, usl_exchange_sums) on the 51.6 MB flat gradient buffer of the Replica workload.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/exp_allreduce.py"""
import importlib, json, os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = f"cuda:{lr}"
dist.init_process_group("nccl", device_id=torch.device(dev))
P = importlib.import_module("uni-slam_b200")
par = importlib.import_module("uni-slam_b200.parallel")
n = 12915392
res = {"world": world, "bytes": n * 4, "lib": os.environ.get("USL_LIB_PATH", "default").split("/")[-1]}


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


bus = lambda us: 2 * (world - 1) / world * n * 4 / (us * 1e-6) / 1e9
x = torch.ones(n, device=dev)
res["nccl_us"] = timeit(lambda: dist.all_reduce(x)); res["nccl_bus_gbs"] = bus(res["nccl_us"])
for mc in (False, True):
    tag = "mc" if mc else "p2p"
    pg = par.PeerGroup(dev, use_multicast=mc)
    g = pg.alloc(n); p = pg.alloc(n)
    if mc and pg.multicast_ptr(g) is None:
        res["mc_unavailable"] = True
        continue
    g.fill_(1.0); torch.cuda.synchronize(); dist.barrier()
    pg.allreduce(g, n); torch.cuda.synchronize()
    res[f"{tag}_correct"] = bool((g == world).all())
    res[f"{tag}_us"] = timeit(lambda: pg.allreduce(g, n)); res[f"{tag}_bus_gbs"] = bus(res[f"{tag}_us"])
    for cap in (2, 4, 16):
        res[f"{tag}_us_cap{cap}"] = timeit(lambda: pg.allreduce(g, n, max_ctas_per_sm=cap))
    fsa = par.FusedShardedAdam(pg, p, g, n, [(0, n, 1e-3)])
    res[f"{tag}_allreduce_adam_us"] = timeit(fsa.step, iters=10)
    if not mc:
        res["barrier_us"] = timeit(pg.barrier)
        acc = torch.ones(16, device=dev)
        res["exchange_sums_us"] = timeit(lambda: pg.exchange_sums(acc))
        small = torch.ones(16, device=dev)
        res["nccl_16_floats_us"] = timeit(lambda: dist.all_reduce(small))
    del fsa, g, p, pg
if rank == 0:
    print(json.dumps(res))
torch.cuda.synchronize(); dist.barrier()
wd = threading.Timer(20.0, lambda: os._exit(0)); wd.daemon = True; wd.start()
dist.destroy_process_group()
