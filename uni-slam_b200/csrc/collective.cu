// Multi-GPU exchange steps of the sharded mapping iteration, hand-written over peer memory: every rank maps
// the other ranks' buffers into its address space (CUDA IPC / symmetric memory, set up by the host layer) and the kernels
// below read and write them directly over NVLink / NVSwitch -- no NCCL call on the critical path.
//
//   usl_exchange_sums      the loss sums / element counts between usl_loss_fwd and the backward: every rank PUSHES its 16
//                          floats into a slot of every peer, raises a flag, waits for the peers' flags and sums the slots.
//                          One tiny kernel (a few microseconds) instead of a latency-bound NCCL all-reduce mid-step.
//   usl_peer_barrier       device-side barrier between the ranks (flag per peer, monotonically increasing epoch).
//   usl_allreduce_sum      two-shot all-reduce of the flat gradient buffer in ONE pass: rank r owns slice r, pulls that
//                          slice from every peer (16-byte loads), sums it once and pushes the SAME bits back into every
//                          peer's buffer (replicas stay in lock-step).  Reads travel in one direction of
//                          the links and writes in the other, so the two shots overlap.
//   usl_allreduce_adam_step  the same pass with the optimiser fused in: the owner of a slice applies Adam to it (state
//                          sharded 1/N per rank, torch.optim.Adam arithmetic of adam.cu) and pushes the NEW PARAMETERS to
//                          every rank instead of the summed gradient: one exchange, Adam / N per rank, no separate
//                          optimiser pass over 12.9 M parameters.
//
// Memory model: data stores to peer memory are followed by __threadfence_system() and a release store of the flag; the
// waiting side reads the flag with an acquire load at system scope.  Kernel boundaries on the caller's stream order the
// barrier kernels against the producers / consumers of the buffers on the same GPU.
#include <cmath>
#include <cstdio>

#include "usl_device.cuh"

namespace usl {

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Peer data is read with ld.global.cg (never from an L1 line of an earlier pass; the owner's L2 is the point of coherence for
// its memory) and written with st.global.cg; ordering against the flags comes from the system-scope fence + release /
// acquire pair of the barrier kernels, so the bulk accesses themselves can stay weak and pipeline freely.
__device__ __forceinline__ float4 ld_peer(const float4 *p) { return __ldcg(p); }
__device__ __forceinline__ void st_peer(float4 *p, const float4 &v) { __stcg(p, v); }

// NVLS: one instruction reads the same address on every GPU of the multicast group and returns the sum, formed inside the
// switch; one store is replicated by the switch to every GPU.  Per GPU and direction the exchange then moves ~(1 + 1/N) x the
// buffer instead of 2 (N-1)/N x: at N = 8, 58 MB instead of 90 MB.
__device__ __forceinline__ float4 mc_ld_reduce(const float4 *p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float4 *p, const float4 &v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// control block of one rank (lives in peer-mapped memory, zero-initialised once): flags written by the peers, the slots of
// the loss-sum exchange (double-buffered by epoch parity) and this rank's own epoch counters.
struct PeerCtrl {
    uint32_t flag_x[USL_MAX_PEERS];                  // exchange: flag_x[p] = epoch of the last push of rank p
    uint32_t flag_b[USL_PEER_CHANNELS][USL_MAX_PEERS];   // barrier:  flag_b[c][p] = epoch of the last arrival of rank p on channel c
    uint32_t epoch_x, epoch_b[USL_PEER_CHANNELS], _pad;  // local counters (only this rank writes them)
    float slots[2][USL_MAX_PEERS][USL_LOSS_SLOTS];   // slots[epoch & 1][p][:] = sums pushed by rank p
};
static_assert(sizeof(PeerCtrl) <= USL_PEER_CTRL_BYTES, "control block larger than the published size");

__global__ void __launch_bounds__(USL_MAX_PEERS * USL_LOSS_SLOTS) exchange_sums_kernel(usl_peers_t P, float *acc) {
    __shared__ uint32_t s_epoch;
    PeerCtrl *me = reinterpret_cast<PeerCtrl *>(P.ctrl[P.rank]);
    const int t = threadIdx.x;
    if (t == 0) { s_epoch = me->epoch_x + 1; me->epoch_x = s_epoch; }
    __syncthreads();
    const uint32_t epoch = s_epoch;
    const int p = t / USL_LOSS_SLOTS, j = t % USL_LOSS_SLOTS;
    if (p < P.world) {
        PeerCtrl *peer = reinterpret_cast<PeerCtrl *>(P.ctrl[p]);
        peer->slots[epoch & 1u][P.rank][j] = acc[j];                  // push (also into my own block)
    }
    __threadfence_system();
    __syncthreads();
    if (t < P.world) {
        st_release_sys(&reinterpret_cast<PeerCtrl *>(P.ctrl[t])->flag_x[P.rank], epoch);
        while (ld_acquire_sys(&me->flag_x[t]) < epoch) { }            // every rank runs the same sequence of exchanges
    }
    __syncthreads();
    if (t < USL_LOSS_SLOTS) {
        float s = 0.f;
        for (int q = 0; q < P.world; ++q) s += me->slots[epoch & 1u][q][t];   // fixed order: the same bits on every rank
        acc[t] = s;
    }
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(usl_peers_t P) {
    PeerCtrl *me = reinterpret_cast<PeerCtrl *>(P.ctrl[P.rank]);
    const int t = threadIdx.x;
    uint32_t epoch = 0;
    const int c = P.channel;                                          // independent barrier sequences for concurrent streams
    if (t == 0) { epoch = me->epoch_b[c] + 1; me->epoch_b[c] = epoch; }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    __threadfence_system();                                           // everything this GPU wrote before is visible system-wide
    if (t < P.world) {
        st_release_sys(&reinterpret_cast<PeerCtrl *>(P.ctrl[t])->flag_b[c][P.rank], epoch);
        while (ld_acquire_sys(&me->flag_b[c][t]) < epoch) { }
    }
}

struct AdamRange {            // [begin, end) in floats of the flat buffer, with its learning rate
    int64_t begin, end;
    float lr;
};
struct ReduceArgs {
    usl_peers_t P;
    int64_t offset;           // first float of the reduced range inside every rank's buffer
    int64_t n4;               // float4 elements in the range
    int64_t slice4;           // float4 elements per rank slice
    // fused optimiser (adam != 0)
    int adam;
    float *param[USL_MAX_PEERS];   // every rank's flat parameter buffer, same layout as the gradient buffer
    float *param_mc;               // multicast mapping of the parameter buffers (NVLS path), or NULL
    float *exp_avg, *exp_avg_sq;   // this rank's state for ITS slice (slice4 * 4 floats each)
    AdamRange range[USL_ADAM_MAX_RANGES];
    int n_ranges;
    float beta1, beta2, eps;
    const int64_t *step_dev;
    int64_t step;
};

__device__ __forceinline__ void adam_update1(float &p, float &m, float &v, float g, float b1, float b2, float eps, float step_size,
                                             float bc2s) {
    m = m + (1.0f - b1) * (g - m);                       // exp_avg.lerp_(grad, 1 - beta1)               (same arithmetic as adam.cu)
    v = v * b2 + (1.0f - b2) * g * g;                    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
    const float denom = sqrtf(v) / bc2s + eps;
    p = p - step_size * (m / denom);
}

template <int WORLD, bool ADAM>
__global__ void __launch_bounds__(256) allreduce_kernel(const __grid_constant__ ReduceArgs A) {
    const int rank = A.P.rank;
    const int64_t lo = (int64_t)rank * A.slice4, hi = min(A.n4, lo + A.slice4);
    __shared__ float s_bc1[USL_ADAM_MAX_RANGES], s_bc2s;
    if (ADAM) {
        if (threadIdx.x == 0) {
            const double t = (double)(A.step_dev ? A.step_dev[0] : A.step);
            for (int r = 0; r < A.n_ranges; ++r) s_bc1[r] = (float)((double)A.range[r].lr / (1.0 - pow((double)A.beta1, t)));
            s_bc2s = (float)sqrt(1.0 - pow((double)A.beta2, t));
        }
        __syncthreads();
    }
    const float4 *src[WORLD];
    float4 *dst[WORLD];
#pragma unroll
    for (int p = 0; p < WORLD; ++p) {
        const int q = (rank + p) % WORLD;                                        // start with the local copy, then walk the ring
        src[p] = reinterpret_cast<const float4 *>(reinterpret_cast<float *>(A.P.buf[q]) + A.offset);
        dst[p] = ADAM ? reinterpret_cast<float4 *>(A.param[q] + A.offset) : reinterpret_cast<float4 *>(reinterpret_cast<float *>(A.P.buf[q]) + A.offset);
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr int U = ADAM ? 1 : (WORLD <= 2 ? 4 : 2);                              // float4 elements per thread and round: loads in flight
    for (int64_t i0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * U) {
        float4 v[U][WORLD];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < hi) {
#pragma unroll
                for (int p = 0; p < WORLD; ++p) v[u][p] = ld_peer(src[p] + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i >= hi) break;
            // a slice is summed once, by its owner, and the SAME bits are pushed to every rank: replicas stay in lock-step
            float4 s = v[u][0];
#pragma unroll
            for (int p = 1; p < WORLD; ++p) { s.x += v[u][p].x; s.y += v[u][p].y; s.z += v[u][p].z; s.w += v[u][p].w; }
            if (ADAM) {
                const int64_t k = (i - lo) * 4;                                      // index into this rank's optimiser state
                float4 m = *reinterpret_cast<float4 *>(A.exp_avg + k), vv = *reinterpret_cast<float4 *>(A.exp_avg_sq + k);
                float4 prm = *reinterpret_cast<const float4 *>(A.param[rank] + A.offset + i * 4);
                float *pe = &prm.x, *me = &m.x, *ve = &vv.x;
                const float *ge = &s.x;
                const int64_t e0 = A.offset + i * 4;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float step_size = 0.f;
                    bool in = false;
                    for (int r = 0; r < A.n_ranges; ++r)
                        if (e0 + c >= A.range[r].begin && e0 + c < A.range[r].end) { step_size = s_bc1[r]; in = true; }
                    if (in) adam_update1(pe[c], me[c], ve[c], ge[c], A.beta1, A.beta2, A.eps, step_size, s_bc2s);
                }
                *reinterpret_cast<float4 *>(A.exp_avg + k) = m;
                *reinterpret_cast<float4 *>(A.exp_avg_sq + k) = vv;
                s = prm;                                                             // what travels to the peers: the new parameters
            }
#pragma unroll
            for (int p = 0; p < WORLD; ++p) st_peer(dst[p] + i, s);
        }
    }
    __threadfence_system();
}

// NVLS form of the same pass: the sum of a slice is formed inside the switch, the result (or the new parameters) is
// replicated by the switch.
template <bool ADAM>
__global__ void __launch_bounds__(256) allreduce_mc_kernel(const __grid_constant__ ReduceArgs A) {
    const int rank = A.P.rank;
    const int64_t lo = (int64_t)rank * A.slice4, hi = min(A.n4, lo + A.slice4);
    __shared__ float s_bc1[USL_ADAM_MAX_RANGES], s_bc2s;
    if (ADAM) {
        if (threadIdx.x == 0) {
            const double t = (double)(A.step_dev ? A.step_dev[0] : A.step);
            for (int r = 0; r < A.n_ranges; ++r) s_bc1[r] = (float)((double)A.range[r].lr / (1.0 - pow((double)A.beta1, t)));
            s_bc2s = (float)sqrt(1.0 - pow((double)A.beta2, t));
        }
        __syncthreads();
    }
    const float4 *src = reinterpret_cast<const float4 *>(reinterpret_cast<float *>(A.P.mc) + A.offset);
    float4 *dst = ADAM ? reinterpret_cast<float4 *>(A.param_mc + A.offset) : reinterpret_cast<float4 *>(reinterpret_cast<float *>(A.P.mc) + A.offset);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
#ifndef USL_MC_UNROLL
#define USL_MC_UNROLL 4
#endif
    constexpr int U = ADAM ? 1 : USL_MC_UNROLL;
    for (int64_t i0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < hi) v[u] = mc_ld_reduce(src + i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i >= hi) break;
            float4 s = v[u];
            if (ADAM) {
                const int64_t k = (i - lo) * 4;
                float4 m = *reinterpret_cast<float4 *>(A.exp_avg + k), vv = *reinterpret_cast<float4 *>(A.exp_avg_sq + k);
                float4 prm = *reinterpret_cast<const float4 *>(A.param[rank] + A.offset + i * 4);
                float *pe = &prm.x, *me = &m.x, *ve = &vv.x;
                const float *ge = &s.x;
                const int64_t e0 = A.offset + i * 4;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float step_size = 0.f;
                    bool in = false;
                    for (int r = 0; r < A.n_ranges; ++r)
                        if (e0 + c >= A.range[r].begin && e0 + c < A.range[r].end) { step_size = s_bc1[r]; in = true; }
                    if (in) adam_update1(pe[c], me[c], ve[c], ge[c], A.beta1, A.beta2, A.eps, step_size, s_bc2s);
                }
                *reinterpret_cast<float4 *>(A.exp_avg + k) = m;
                *reinterpret_cast<float4 *>(A.exp_avg_sq + k) = vv;
                s = prm;
            }
            mc_st(dst + i, s);
        }
    }
    __threadfence_system();
}

static int check_peers(const usl_peers_t *P, const char *who) {
    if (!P || P->world < 1 || P->world > USL_MAX_PEERS || P->rank < 0 || P->rank >= P->world || P->channel < 0 || P->channel >= USL_PEER_CHANNELS) {
        set_error("%s: bad peer group", who); return 1;
    }
    for (int p = 0; p < P->world; ++p)
        if (!P->ctrl[p]) { set_error("%s: peer %d has no control block", who, p); return 1; }
    return 0;
}

template <bool ADAM>
static int launch_reduce(const ReduceArgs &A, cudaStream_t s) {
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int64_t mine = A.slice4;
    int64_t blocks = (mine + 255) / 256;
    const bool use_mc = A.P.mc != nullptr && A.P.world > 1 && (!ADAM || A.param_mc != nullptr);
    // grid caps as measured at 51.6 MB (tools/exp_allreduce.py): multimem likes many CTAs (16/SM: 160 us at N = 8, 8/SM: 168);
    // the peer-to-peer form at 8 ranks likes few (2/SM: 170 us, 8/SM: 190, 16/SM: 206) -- seven remote streams per thread
    const int per_sm = A.P.max_ctas_per_sm > 0 ? A.P.max_ctas_per_sm : use_mc ? 16 : A.P.world >= 8 ? 2 : 8;
    const int64_t cap = (int64_t)n_sm * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (use_mc) {
        allreduce_mc_kernel<ADAM><<<(unsigned)blocks, 256, 0, s>>>(A);
        return check_launch(ADAM ? "usl_allreduce_adam_step (multimem)" : "usl_allreduce_sum (multimem)");
    }
    switch (A.P.world) {
#define USL_CASE(W) case W: allreduce_kernel<W, ADAM><<<(unsigned)blocks, 256, 0, s>>>(A); break;
        USL_CASE(1) USL_CASE(2) USL_CASE(3) USL_CASE(4) USL_CASE(5) USL_CASE(6) USL_CASE(7) USL_CASE(8)
#undef USL_CASE
        default: set_error("usl_allreduce: world size %d not supported", A.P.world); return 1;
    }
    return check_launch(ADAM ? "usl_allreduce_adam_step" : "usl_allreduce_sum");
}

}  // namespace usl

using namespace usl;

extern "C" {

int usl_peer_ctrl_bytes(void) { return USL_PEER_CTRL_BYTES; }

int usl_exchange_sums(const usl_peers_t *P, float *acc, usl_stream_t stream) {
    if (check_peers(P, "usl_exchange_sums")) return 1;
    if (!acc) { set_error("usl_exchange_sums: null accumulator"); return 1; }
    exchange_sums_kernel<<<1, USL_MAX_PEERS * USL_LOSS_SLOTS, 0, (cudaStream_t)stream>>>(*P, acc);
    return check_launch("usl_exchange_sums");
}

int usl_peer_barrier(const usl_peers_t *P, usl_stream_t stream) {
    if (check_peers(P, "usl_peer_barrier")) return 1;
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*P);
    return check_launch("usl_peer_barrier");
}

static int fill_reduce(ReduceArgs &A, const usl_peers_t *P, int64_t offset_floats, int64_t n_floats, const char *who) {
    if (check_peers(P, who)) return 1;
    if ((offset_floats & 3) || (n_floats & 3) || n_floats < 0) { set_error("%s: offset and length must be multiples of 4 floats", who); return 1; }
    for (int p = 0; p < P->world; ++p)
        if (!P->buf[p] || ((uintptr_t)P->buf[p] & 15u)) { set_error("%s: peer %d buffer missing or not 16-byte aligned", who, p); return 1; }
    A.P = *P; A.offset = offset_floats; A.n4 = n_floats / 4;
    A.slice4 = (A.n4 + P->world - 1) / P->world;
    A.adam = 0; A.n_ranges = 0; A.param_mc = nullptr;
    return 0;
}

int usl_allreduce_sum(const usl_peers_t *P, int64_t offset_floats, int64_t n_floats, usl_stream_t stream) {
    ReduceArgs A;
    if (fill_reduce(A, P, offset_floats, n_floats, "usl_allreduce_sum")) return 1;
    if (n_floats == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    peer_barrier_kernel<<<1, 32, 0, s>>>(*P);            // every rank's gradients are complete
    if (launch_reduce<false>(A, s)) return 1;
    peer_barrier_kernel<<<1, 32, 0, s>>>(*P);            // every slice has landed everywhere
    return check_launch("usl_allreduce_sum (barrier)");
}

int usl_allreduce_adam_slice_floats(int world, int64_t n_floats, int64_t *slice_floats) {
    if (world < 1 || world > USL_MAX_PEERS || !slice_floats || (n_floats & 3)) { set_error("usl_allreduce_adam_slice_floats: bad arguments"); return 1; }
    *slice_floats = ((n_floats / 4 + world - 1) / world) * 4;
    return 0;
}

int usl_allreduce_adam_step(const usl_peers_t *P, float *const *param, float *param_mc, int64_t offset_floats, int64_t n_floats,
                            float *exp_avg, float *exp_avg_sq, const usl_adam_range_t *ranges, int n_ranges, float beta1,
                            float beta2, float eps, int64_t step, const int64_t *step_dev, usl_stream_t stream) {
    ReduceArgs A;
    if (fill_reduce(A, P, offset_floats, n_floats, "usl_allreduce_adam_step")) return 1;
    if (!param || !exp_avg || !exp_avg_sq || !ranges || n_ranges < 1 || n_ranges > USL_ADAM_MAX_RANGES || (step < 1 && !step_dev)) {
        set_error("usl_allreduce_adam_step: bad arguments (1..%d learning-rate ranges, step >= 1)", USL_ADAM_MAX_RANGES); return 1;
    }
    for (int p = 0; p < P->world; ++p) {
        if (!param[p] || ((uintptr_t)param[p] & 15u)) { set_error("usl_allreduce_adam_step: peer %d parameter buffer missing or unaligned", p); return 1; }
        A.param[p] = param[p];
    }
    A.param_mc = param_mc;
    A.adam = 1; A.exp_avg = exp_avg; A.exp_avg_sq = exp_avg_sq; A.n_ranges = n_ranges;
    for (int r = 0; r < n_ranges; ++r) { A.range[r].begin = ranges[r].begin; A.range[r].end = ranges[r].end; A.range[r].lr = ranges[r].lr; }
    A.beta1 = beta1; A.beta2 = beta2; A.eps = eps; A.step = step; A.step_dev = step_dev;
    if (n_floats == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    peer_barrier_kernel<<<1, 32, 0, s>>>(*P);
    if (launch_reduce<true>(A, s)) return 1;
    peer_barrier_kernel<<<1, 32, 0, s>>>(*P);
    return check_launch("usl_allreduce_adam_step (barrier)");
}

}  // extern "C"
