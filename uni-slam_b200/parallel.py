"""Host-side multi-GPU helpers (S§8e): one process per GPU, torch.distributed for the plumbing.

* mapping: ray batches shard across ranks (every rank draws its own batch); two exchange steps per iteration --
  the 12 loss sums/counts between ``usl_loss_fwd`` and ``usl_loss_bwd`` (so every mean divides by the GLOBAL
  element count), and one all-reduce of the flat gradient buffer [tables | decoders | beta] (+ pose gradients)
  after the backward.  Identical Adam on every rank then keeps the replicas in lock-step.
* dense SDF query: y-slabs of the (ny, nx, nz) volume, no data-path collective, slabs gathered to rank 0.
* tracking: replicas only (single GPU, BASELINE.json north_star).
"""
from typing import Tuple

import torch
import torch.distributed as dist


def slab_range(ny: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous y-slab [begin, end) of rank `rank`; slabs differ by at most one row and tile [0, ny)."""
    base, rem = divmod(ny, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def attach_mapping_collectives(step, group=None, overlap: bool = False):
    """Wire the exchange steps of a sharded mapping iteration onto a MappingStep.  Returns reduce_rest(), to be called
    after step.run().  With overlap=True the colour-table gradient (87 % of the bytes) is all-reduced on a side stream
    while the sdf half of usl_field_bwd still runs (CUDA-graph capturable).  Measured on 2 B200: SLOWER (0.80 vs 0.75 ms per
    step) -- the two half launches lose the sdf/colour CTA interleaving and NCCL's reduction competes with the atomics for
    L2 -- so the default is one all-reduce of the flat gradient buffer after the backward."""
    def acc_hook(acc):
        dist.all_reduce(acc, group=group)
    step.acc_hook = acc_hook
    fs = step.fs
    if not overlap:
        def reduce_all():
            dist.all_reduce(fs.g_grads, group=group)          # tables + decoders + beta + pose gradients: one collective
        return reduce_all
    side = torch.cuda.Stream()

    def rgb_hook(g_rgb):
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            dist.all_reduce(g_rgb, group=group)
    step.rgb_grads_hook = rgb_hook

    def reduce_rest():
        dist.all_reduce(fs.g_sdf_table, group=group)
        dist.all_reduce(fs.g_flat, group=group)
        torch.cuda.current_stream().wait_stream(side)
    return reduce_rest


def finalize_loss(acc: torch.Tensor, w_fs, w_center, w_tail, w_depth, w_color) -> torch.Tensor:
    """Host mirror of usl_loss_finalize for logging: loss from the (all-reduced) sums/counts of usl_loss_fwd.
    Slots: 0 fs 1 center 2 tail 3 depth 4 colour sums; 5 n_front 6 n_center 7 n_tail 8 n_mask 10 n_colour_terms."""
    return (w_fs * acc[0] / acc[5] + w_center * acc[1] / acc[6] + w_tail * acc[2] / acc[7]
            + w_color * acc[4] / acc[10] + w_depth * acc[3] / acc[8])


def gather_slabs(local: torch.Tensor, ny: int, nx: int, nz: int, rank: int, world: int, group=None):
    """Gather the per-rank (rows*nx*nz,) SDF slabs into the full (ny*nx*nz,) volume on rank 0 (None elsewhere)."""
    sizes = [(slab_range(ny, r, world)[1] - slab_range(ny, r, world)[0]) * nx * nz for r in range(world)]
    if world == 1:
        return local
    # dist.gather needs equal sizes: pad every slab to the largest one, trim on rank 0
    mx = max(sizes)
    padded = local if local.numel() == mx else torch.cat([local, local.new_zeros(mx - local.numel())])
    bufs = [torch.empty(mx, dtype=local.dtype, device=local.device) for _ in sizes] if rank == 0 else None
    dist.gather(padded, bufs, dst=0, group=group)
    return torch.cat([b[:s_] for b, s_ in zip(bufs, sizes)]) if rank == 0 else None
