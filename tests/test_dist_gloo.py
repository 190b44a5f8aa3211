"""CPU, world_size 2, gloo: the host-side multi-GPU logic (S§8e) -- sharded loss normalisation with all-reduced
sums/counts equals the single-process loss; y-slab partition and gather reproduce the full volume."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _loss_acc(sdf, z, gt, depth, rgb, gt_color, punc, tr):
    """CPU restatement of usl_loss_fwd's sums/counts (mapping mode) used as the per-rank partial result."""
    m = (gt > 0) & ((1 - punc) > 0.99)
    zz, ss, gg = z[m], sdf[m], gt[m][:, None]
    front = zz < gg - tr; back = zz > gg + tr
    center = (zz > gg - 0.4 * tr) & (zz < gg + 0.4 * tr); tail = ~front & ~back & ~center
    e = (zz + ss * tr) - gg
    acc = torch.zeros(16, dtype=torch.float64)
    acc[0] = ((ss - 1) ** 2)[front].sum(); acc[1] = (e ** 2)[center].sum(); acc[2] = (e ** 2)[tail].sum()
    acc[3] = ((gt - depth) ** 2)[m].sum(); acc[4] = ((gt_color - rgb) ** 2).sum()
    acc[5], acc[6], acc[7], acc[8] = front.sum(), center.sum(), tail.sum(), m.sum()
    acc[9] = gt.numel(); acc[10] = 3 * gt.numel()
    return acc


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    import importlib
    par = importlib.import_module("uni-slam_b200.parallel")
    from oracle import path_ref
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)                       # same data on every rank; each takes its shard
    R, S, tr = 64, 40, 0.06
    gt = torch.rand(R, generator=g, dtype=torch.float64) * 3 + 0.5
    gt[::9] = 0
    z = torch.sort(torch.rand(R, S, generator=g, dtype=torch.float64) * 1.2 * gt.clamp_min(1)[:, None], -1)[0]
    sdf = torch.tanh((gt.clamp_min(1)[:, None] - z) * 5)
    depth = gt + 0.01 * torch.randn(R, generator=g, dtype=torch.float64)
    rgb = torch.rand(R, 3, generator=g, dtype=torch.float64); gtc = torch.rand(R, 3, generator=g, dtype=torch.float64)
    punc = torch.rand(R, generator=g, dtype=torch.float64) * 0.012
    lo, hi = rank * R // world, (rank + 1) * R // world
    acc = _loss_acc(sdf[lo:hi], z[lo:hi], gt[lo:hi], depth[lo:hi], rgb[lo:hi], gtc[lo:hi], punc[lo:hi], tr)
    dist.all_reduce(acc)                                        # exchange step 1 of the sharded mapping iteration
    loss = par.finalize_loss(acc, 5.0, 200.0, 10.0, 0.1, 5.0)
    ret = (None, punc, depth, rgb, sdf, z, None)
    ref = path_ref.mapping_loss(ret, gt, gtc, tr)
    ok_loss = abs(float(loss) - float(ref)) < 1e-9 * abs(float(ref))
    # y-slab partition + gather
    ny, nx, nz = 7, 3, 2
    full = torch.arange(ny * nx * nz, dtype=torch.float32)
    b, e = par.slab_range(ny, rank, world)
    out = par.gather_slabs(full[b * nx * nz:e * nx * nz].clone(), ny, nx, nz, rank, world)
    ok_slab = True if rank != 0 else bool(torch.equal(out, full))
    # image rows sharded the same way (frame renderer): an (H,W,3) colour image in row slabs of pixel ranges
    H, W = 5, 4
    img = torch.arange(H * W * 3, dtype=torch.float32)
    r0, r1 = par.slab_range(H, rank, world)
    p0, p1 = r0 * W, r1 * W                                     # the pixel range RenderImageStep.run(pixel_begin, pixel_end) takes
    out = par.gather_slabs(img[p0 * 3:p1 * 3].clone(), H, W, 3, rank, world)
    ok_slab = ok_slab and (True if rank != 0 else bool(torch.equal(out, img)))
    q.put((rank, ok_loss, ok_slab))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_loss_and_slabs():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok_loss and ok_slab for _, ok_loss, ok_slab in res), res


def test_slab_ranges_tile_the_volume():
    import importlib
    par = importlib.import_module("uni-slam_b200.parallel")
    for ny in (1, 7, 510, 511):
        for world in (1, 2, 4, 8):
            r = [par.slab_range(ny, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == ny and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
