"""Host-side multi-GPU helpers (S§8e): one process per GPU, torch.distributed for the plumbing.

* mapping: ray batches shard across ranks (every rank draws its own batch); two exchange steps per iteration --
  the 12 loss sums/counts between ``usl_loss_fwd`` and ``usl_loss_bwd`` (so every mean divides by the GLOBAL
  element count), and one all-reduce of the flat gradient buffer [tables | decoders | beta] (+ pose gradients)
  after the backward.  Identical Adam on every rank then keeps the replicas in lock-step.
* dense SDF query: y-slabs of the (ny, nx, nz) volume, no data-path collective, slabs gathered to rank 0.
* tracking: replicas only (single GPU, BASELINE.json north_star).
"""
from ctypes import byref, c_int64, c_void_p
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib as L


def slab_range(ny: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous y-slab [begin, end) of rank `rank`; slabs differ by at most one row and tile [0, ny)."""
    base, rem = divmod(ny, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class PeerGroup:
    """Peer-mapped (symmetric) memory for the hand-written exchange kernels of csrc/collective.cu.

    torch.distributed's symmetric-memory allocator is the plumbing: one allocation + rendezvous per buffer maps every
    rank's copy into every rank's address space (CUDA IPC over NVLink / NVSwitch).  The kernels then read and write the
    peers' buffers directly; NCCL is not called on the iteration's critical path.

      alloc(numel)          zero-filled fp32 symmetric tensor (pass as MappingStep(grad_alloc=pg.alloc): the flat gradient buffer)
      exchange_sums(acc)    usl_exchange_sums: the 16 loss sums / counts become global (between loss_fwd and the backward)
      allreduce(t, n)       usl_allreduce_sum over the first n floats of a tensor obtained from alloc()
      FusedShardedAdam      usl_allreduce_adam_step (see below)

    use_multicast=True runs the reductions through the NVSwitch (multimem.ld_reduce / multimem.st on the allocator's
    multicast mapping).  Measured on this pool's B200 boxes for the 51.6 MB gradient buffer (us; NCCL / peer-to-peer /
    multimem): N = 2: 120 / 111 / 176, N = 4: 156 / 155 / 169, N = 8: 232 / 190 / 168 -- so "auto" uses peer-to-peer loads and
    stores below 8 ranks and the switch-side reduction from 8."""

    def __init__(self, device, group=None, use_multicast="auto"):
        import torch.distributed._symmetric_memory as symm_mem
        self._sm = symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.gname = self.group.group_name
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise RuntimeError("PeerGroup: at most 8 ranks (one NVSwitch domain)")
        self.device = device
        self._handles = {}
        # "auto": for a stand-alone exchange the switch-side reduction pays off from 8 ranks (measured, see the class docstring);
        # an exchange that runs BESIDE the backward (attach_peer_collectives) uses it at every world size
        self.multicast_mode = use_multicast
        self.use_multicast = (self.world >= 8) if use_multicast == "auto" else bool(use_multicast)
        nb = L.load().usl_peer_ctrl_bytes()
        self.ctrl = self._symm_zeros(nb // 4)
        self._ctrl_ptrs = self._handles[self.ctrl.data_ptr()].buffer_ptrs

    @classmethod
    def single(cls, device):
        """World of one rank without torch.distributed: the 'peers' are this process's own buffers.  The kernels run
        unchanged (used by the single-GPU tests of collective.cu and by callers that want one code path for N = 1)."""
        self = cls.__new__(cls)
        self._sm, self.group, self.gname, self.rank, self.world, self.device = None, None, None, 0, 1, device
        self._handles = {}
        self.use_multicast = False; self.multicast_mode = False
        self.ctrl = self._symm_zeros(L.load().usl_peer_ctrl_bytes() // 4)
        self._ctrl_ptrs = self._handles[self.ctrl.data_ptr()].buffer_ptrs
        return self

    def _symm_zeros(self, numel):
        if self._sm is None:                               # single(): plain device memory
            t = torch.zeros(int(numel), dtype=torch.float32, device=self.device)
            self._handles[t.data_ptr()] = type("LocalHandle", (), {"buffer_ptrs": [t.data_ptr()]})()
            return t
        t = self._sm.empty(int(numel), dtype=torch.float32, device=self.device)
        h = self._sm.rendezvous(t, group=self.gname)
        t.zero_()
        torch.cuda.synchronize()
        dist.barrier(self.group)                       # nobody touches a peer's buffer before it is zeroed
        self._handles[t.data_ptr()] = h
        return t

    def alloc(self, numel):
        return self._symm_zeros(numel)

    def peers(self, t: torch.Tensor, multicast=None) -> L.Peers:
        """usl_peers_t for a tensor obtained from alloc(): every rank's copy of it + the control blocks.
        multicast: None = the group's default for stand-alone exchanges; True / False = this call's choice."""
        h = self._handles.get(t.data_ptr())
        if h is None:
            raise ValueError("PeerGroup.peers: tensor was not allocated by this PeerGroup")
        P = L.Peers()
        P.rank, P.world = self.rank, self.world
        for p in range(self.world):
            P.buf[p] = h.buffer_ptrs[p]
            P.ctrl[p] = self._ctrl_ptrs[p]
        P.mc = self.multicast_ptr(t) if (self.use_multicast if multicast is None else multicast) else None
        return P

    def multicast_ptr(self, t: torch.Tensor):
        """NVLS multicast mapping of a buffer from alloc() (None when the fabric / allocator offers none)."""
        h = self._handles.get(t.data_ptr())
        mc = getattr(h, "multicast_ptr", 0) if h is not None else 0
        return int(mc) if mc else None

    def exchange_sums(self, acc: torch.Tensor):
        P = self.peers(self.ctrl)
        L.call("usl_exchange_sums", byref(P), L.ptr(acc), L.stream())

    def allreduce(self, t: torch.Tensor, n_floats: int, offset_floats: int = 0, channel: int = 0, max_ctas_per_sm: int = 0, multicast=None):
        """channel: exchanges in flight at the same time (different streams) need different barrier channels;
        max_ctas_per_sm=1: a small grid that shares the SMs with a compute kernel (the pass is bound by the links)."""
        P = self.peers(t, multicast)
        P.channel, P.max_ctas_per_sm = int(channel), int(max_ctas_per_sm)
        L.call("usl_allreduce_sum", byref(P), int(offset_floats), int(n_floats), L.stream())

    def barrier(self):
        P = self.peers(self.ctrl)
        L.call("usl_peer_barrier", byref(P), L.stream())


class FusedShardedAdam:
    """usl_allreduce_adam_step: gradient reduction, Adam and parameter broadcast in one pass over peer memory.

    ``params`` and ``grads`` are flat symmetric buffers of the same layout (PeerGroup.alloc); ``ranges`` = [(begin, end, lr)]
    in floats gives the learning rate of each stretch of the layout (the groups of Mapper.create_optimizer,
    src/Mapper.py:111-139); rank r owns slice r of the buffers and keeps Adam state for that slice only."""

    def __init__(self, pg: PeerGroup, params: torch.Tensor, grads: torch.Tensor, n_floats: int, ranges: Sequence[Tuple[int, int, float]],
                 betas=(0.9, 0.999), eps=1e-8):
        self.pg, self.params, self.grads, self.n = pg, params, grads, int(n_floats)
        sl = c_int64(0)
        L.call("usl_allreduce_adam_slice_floats", pg.world, self.n, byref(sl))
        self.exp_avg = torch.zeros(sl.value, device=params.device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(sl.value, device=params.device, dtype=torch.float32)
        self.ranges = (L.AdamRange * len(ranges))()
        for a, (b, e, lr) in zip(self.ranges, ranges):
            a.begin, a.end, a.lr = int(b), int(e), float(lr)
        self.betas, self.eps = betas, eps
        self.step_dev = torch.zeros((1,), device=params.device, dtype=torch.int64)
        hp = pg._handles[params.data_ptr()]
        self._pptrs = (c_void_p * 8)(*[hp.buffer_ptrs[p] for p in range(pg.world)] + [None] * (8 - pg.world))

    @torch.no_grad()
    def step(self):
        self.step_dev += 1                                            # device-side step count: CUDA-graph replayable
        P = self.pg.peers(self.grads)
        L.call("usl_allreduce_adam_step", byref(P), self._pptrs, self.pg.multicast_ptr(self.params) if self.pg.use_multicast else None, 0, self.n, L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq), self.ranges,
               len(self.ranges), self.betas[0], self.betas[1], self.eps, 0, L.ptr(self.step_dev), L.stream())


def attach_peer_collectives(step, pg: PeerGroup, overlap=None, side_ctas_per_sm: int = 1):
    """Wire the hand-written exchange steps onto a MappingStep whose gradient buffer came from pg.alloc.  Returns
    reduce_grads(), to be called after step.run().

    overlap=False: one usl_allreduce_sum over the whole buffer after the backward.
    overlap=True: the backward runs as two launches, colour grid first; the colour-table gradient (87 % of the bytes, the
    first contiguous range of the flat buffer) is exchanged on a side stream by a one-CTA-per-SM kernel WHILE the sdf half
    of the backward runs (which leaves one CTA slot per SM free for it); reduce_grads() then exchanges the remaining range
    [sdf table | decoders | beta | poses] and joins the side stream.  CUDA-graph capturable (fork / join on events).
    overlap=None (default): on, with both exchanges going through the switch (multimem), whenever the fabric offers a multicast
    mapping and the group was not built with use_multicast=False; off otherwise.  Measured:
      * peer-to-peer loads / stores (kernel timeline, bench.py --trace): the exchange does run beside the sdf half, but its
        loads and the scatter's atomics share the SMs' LSU path -- it stretches from 90 to 150 us and the split backward costs
        20 us more than the single launch: 565 vs 577 us per step at N = 2, 623 vs 613 us at N = 4 -- no gain;
      * multimem: the switch does the arithmetic, one CTA per SM keeps the links busy, and the exchange hides almost
        completely behind the sdf half: N = 8: 614 -> 540 us per step (1 CTA per SM; 547 / 545 us with 2 / 4); N = 2: 559
        (peer-to-peer, not overlapped) -> 543 us, although the stand-alone multimem exchange is the slower one there (639 us
        per step when it is not overlapped)."""
    fs = step.fs
    mc_ok = pg.world > 1 and pg.multicast_mode is not False and pg.multicast_ptr(fs.g_all) is not None
    if overlap is None:
        overlap = mc_ok
    mc = True if (overlap and mc_ok) else None              # beside the backward: through the switch at every world size
    step.acc_hook = pg.exchange_sums
    fs = step.fs
    if not overlap:
        def reduce_all():
            pg.allreduce(fs.g_all, fs.n_grad_padded)
        return reduce_all
    side = torch.cuda.Stream()
    n_rgb = fs.g_rgb_table.numel()
    assert fs.g_rgb_table.data_ptr() == fs.g_all.data_ptr() and n_rgb % 4 == 0

    def rgb_hook(_g_rgb):
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            pg.allreduce(fs.g_all, n_rgb, 0, channel=1, max_ctas_per_sm=side_ctas_per_sm, multicast=mc)
    step.rgb_grads_hook = rgb_hook
    step.bwd_leave_room = True

    def reduce_rest():
        pg.allreduce(fs.g_all, fs.n_grad_padded - n_rgb, n_rgb, channel=0, multicast=mc)
        torch.cuda.current_stream().wait_stream(side)
    return reduce_rest


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
