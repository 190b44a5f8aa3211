"""Experiment (torchrun): NCCL all_reduce vs torch symmetric-memory all-reduce kernels on the 51.6 MB gradient buffer."""
import os, sys, time, json
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = f"cuda:{lr}"
dist.init_process_group("nccl", device_id=torch.device(dev))
n = 12913696 + 1544
res = {}
def timeit(fn, iters=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
x = torch.ones(n, device=dev)
res["nccl_us"] = timeit(lambda: dist.all_reduce(x))
try:
    import torch.distributed._symmetric_memory as symm_mem
    gname = dist.group.WORLD.group_name
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    symm_mem.rendezvous(t, group=gname)
    t.fill_(1.0)
    for name in ("two_shot_all_reduce_", "multimem_all_reduce_"):
        try:
            op = getattr(torch.ops.symm_mem, name)
            t.fill_(1.0); torch.cuda.synchronize(); dist.barrier()
            op(t, "sum", gname); torch.cuda.synchronize()
            ok = bool((t[:1000] == world).all())
            res[name + "_correct"] = ok
            res[name + "_us"] = timeit(lambda: op(t, "sum", gname))
        except Exception as e:
            res[name + "_err"] = f"{type(e).__name__}: {str(e)[:120]}"
except Exception as e:
    res["symm_err"] = f"{type(e).__name__}: {str(e)[:200]}"
if rank == 0: print(json.dumps(res))
dist.barrier(); torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
