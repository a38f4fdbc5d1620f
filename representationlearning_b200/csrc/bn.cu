// Training-mode BatchNorm / SyncBatchNorm over NHWC activations, fused with the activation that
// follows it in the reference (ReLU in HRNet: _hrnet_rssformer.py:226-245; exact-erf GELU in the
// FFN: modules/ffn_block.py:246-263) and with the residual add of BasicBlock/Bottleneck.
//
//   stats    : per-block shifted (mean, M2) partials per channel          -> rss_bn_stats
//   combine  : Chan-combine partials (also across ranks for SyncBN)       -> rss_bn_combine
//   finalize : scale/shift, saved mean/invstd, running-stat update        -> rss_bn_finalize
//   apply    : y = act(x*scale + shift [+ residual])                      -> rss_bn_act_fwd
//   backward : dz = dy*act'(.), sums (sum dz, sum dz*xhat)                -> rss_bn_bwd_reduce
//              dx = gamma*invstd*(dz - sum_dz/n - xhat*sum_dzxhat/n)      -> rss_bn_bwd_apply
// All kernels are HBM-bound: each thread owns 8 consecutive channels (one 16/32-byte vector).
#include <stdlib.h>
#include "common.cuh"

namespace rss {

struct BnGeom { int cg; int rpb; int threads; };
static inline BnGeom bn_geom(int C) {
    BnGeom g; g.cg = C / 8; g.rpb = 256 / g.cg; if (g.rpb < 1) g.rpb = 1; g.threads = g.cg * g.rpb; return g;
}
// Grid sizes of the BatchNorm kernels: resident 256-thread blocks per SM.  The step is a multi-stream DAG of ~2800 short kernels
// (average concurrency 1.8): a grid of 8 blocks per SM takes every thread slot, and while it runs the kernels the other streams
// have ready cannot become resident.  Smaller grids cost a little per kernel and win for the step.  Measured on the B=16 step
// (img/s, each row on one box, back to back; round 2, after the stem / tail kernels):
//   apply = big, ticket 2:   8 -> 532.0 / 534.8   6 -> 533.3   4 -> 541.7   3 -> 539.8   2 -> 537.5   16 -> 526.9
//   apply 4, ticket 2, big:  8 -> 537.6           4 -> 542.2   3 -> 540.9   2 -> 539.9
//   apply 4, big 4, ticket:  3 -> 541.9           2 -> 542.2   1 -> 548.4
//   ticket 1, big 4, apply:  4 -> 545.3           3 -> 545.6   2 -> 549.5
// (round 1, with the library stem and the old tail, had found ticket 1 -> 457, 2 -> 465, 4 -> 459, 8 -> 456.)
//
// RSS_BN_TICKET_BPSM (default 1): kernels that end in a device-wide ticket (statistics, backward reduce) on tensors < 24 MB: every
// block pays one same-address atomic round trip at its end, so fewer, fatter blocks also shorten the tail.
static inline int bn_ticket_bpsm() {
    static int v = 0;
    if (v == 0) { const char* e = getenv("RSS_BN_TICKET_BPSM"); v = e ? atoi(e) : 1; if (v < 1 || v > 8) v = 1; }
    return v;
}
// RSS_BN_APPLY_BPSM (default 2): the streaming apply kernels on tensors < 24 MB (L2-resident layers of the multi-stream part)
static inline int bn_apply_bpsm() {
    static int v = 0;
    if (v == 0) { const char* e = getenv("RSS_BN_APPLY_BPSM"); v = e ? atoi(e) : 2; if (v < 1 || v > 16) v = 2; }
    return v;
}
// RSS_BN_BIG_BPSM (default 4): every BatchNorm kernel on tensors >= 24 MB (stem, layer1, the FFN's 67 MB hidden activations, the
// neck): HBM-bound passes that need bytes in flight -- 2 blocks per SM kept only ~32 KB of loads in flight per SM (37 us for a
// 67 MB read) -- but even these lose at 8 (the weight-gradient streams run beside them)
static inline int bn_big_bpsm() {
    static int v = 0;
    if (v == 0) { const char* e = getenv("RSS_BN_BIG_BPSM"); v = e ? atoi(e) : 4; if (v < 1 || v > 16) v = 4; }
    return v;
}
static inline bool bn_is_big(int64_t rows, int C, int dtype) { return rows * C * (dtype == RSS_F32 ? 4 : 2) >= ((int64_t)24 << 20); }
static inline int bn_ticket_bpsm_for(int64_t rows, int C, int dtype) { return bn_is_big(rows, C, dtype) ? bn_big_bpsm() : bn_ticket_bpsm(); }
static inline int bn_apply_bpsm_for(int64_t rows, int C, int dtype) { return bn_is_big(rows, C, dtype) ? bn_big_bpsm() : bn_apply_bpsm(); }
static inline int bn_grid(int64_t rows, int rpb, int per_sm) {
    int64_t g = (rows + rpb - 1) / rpb;
    const int64_t cap = (int64_t)num_sms() * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void bn_block_partial(const T* __restrict__ x, float* __restrict__ part /*[grid][C][2]*/,
                                                 float* __restrict__ cnt /*[grid]*/, int64_t rows, int C, int cg, int rpb, float* sm) {
    const int sub = threadIdx.x % cg, r = threadIdx.x / cg;
    float K[8], s[8], q[8];
    int64_t n = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { K[i] = 0.f; s[i] = 0.f; q[i] = 0.f; }
    // contiguous chunk of rows per block keeps the shift K representative
    const int64_t chunk = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * chunk;
    const int64_t r1 = (r0 + chunk < rows) ? r0 + chunk : rows;
    for (int64_t row = r0 + r; row < r1; row += rpb) {
        float v[8];
        load8(x + row * C + sub * 8, v);
        if (n == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) K[i] = v[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[i] - K[i]; s[i] += d; q[i] += d * d; }
        ++n;
    }
    const float fn = (float)n;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float* e = sm + ((size_t)r * C + sub * 8 + i) * 3;
        const float md = n ? s[i] / fn : 0.f;
        e[0] = fn; e[1] = K[i] + md; e[2] = n ? fmaxf(q[i] - s[i] * md, 0.f) : 0.f;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float na = 0.f, ma = 0.f, qa = 0.f;
        for (int rr = 0; rr < rpb; ++rr) {
            const float* e = sm + ((size_t)rr * C + c) * 3;
            const float nb = e[0];
            if (nb > 0.f) {
                const float nn = na + nb, d = e[1] - ma;
                ma += d * (nb / nn);
                qa += e[2] + d * d * (na * nb / nn);
                na = nn;
            }
        }
        part[((size_t)blockIdx.x * C + c) * 2] = ma;
        part[((size_t)blockIdx.x * C + c) * 2 + 1] = qa;
        if (c == 0) cnt[blockIdx.x] = na;
    }
}

template <typename T>
__global__ void bn_stats_kernel(const T* __restrict__ x, float* __restrict__ part, float* __restrict__ cnt,
                                int64_t rows, int C, int cg, int rpb) {
    extern __shared__ float sm[];           // [rpb][C][3]  (n, mean, M2)
    bn_block_partial<T>(x, part, cnt, rows, C, cg, rpb, sm);
}

// Chan-merge of the partials of one channel by one warp (lanes stride over partials, shuffle tree at the end)
__device__ __forceinline__ void bn_warp_combine(const float* __restrict__ part, const float* __restrict__ cnt, int nparts, int C,
                                                int c, int lane, float& na, float& ma, float& qa) {
    na = 0.f; ma = 0.f; qa = 0.f;
    for (int p = lane; p < nparts; p += 32) {
        const float nb = cnt[p];
        if (nb > 0.f) {
            const float mb = part[((size_t)p * C + c) * 2], qb = part[((size_t)p * C + c) * 2 + 1];
            const float nn = na + nb, d = mb - ma;
            ma += d * (nb / nn);
            qa += qb + d * d * (na * nb / nn);
            na = nn;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float nb = __shfl_xor_sync(0xffffffffu, na, o), mb = __shfl_xor_sync(0xffffffffu, ma, o),
                    qb = __shfl_xor_sync(0xffffffffu, qa, o);
        const float nn = na + nb;
        if (nn > 0.f) {
            const float d = mb - ma;
            const float m = (na * ma + nb * mb) / nn;       // symmetric: both lanes of a pair get the same triple
            qa = qa + qb + d * d * (na * nb / nn);
            ma = m;
            na = nn;
        }
    }
}

__device__ __forceinline__ void bn_finalize_channel(float n, float mean, float m2, int c, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, float* __restrict__ running_mean,
                                                    float* __restrict__ running_var, float momentum, float eps,
                                                    float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                                    float* __restrict__ scale, float* __restrict__ shift,
                                                    const float* __restrict__ pre_bias = nullptr) {
    const float invstd = rsqrtf(m2 / n + eps);                    // biased variance normalises
    mean_out[c] = mean;
    invstd_out[c] = invstd;
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - mean * sc;
    if (running_mean) {
        // pre_bias: the producing conv's bias was NOT added to x (a per-channel shift cancels in the normalisation); the
        // running mean is the only place where it is visible
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (mean + (pre_bias ? pre_bias[c] : 0.f));
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (m2 / fmaxf(n - 1.f, 1.f));   // unbiased estimate
    }
}

// single-GPU training path: statistics and finalize in ONE launch.  Per-channel sums of d = x - K and d^2 (K = the running
// mean: a shift close to the batch mean keeps E[d^2] - E[d]^2 well conditioned in fp32) are reduced per block and added
// atomically to a persistent zeroed scratch; the last block to arrive (device-wide ticket) turns them into
// scale/shift + running statistics and clears the scratch for the next launch on the stream.
template <typename T>
__global__ void bn_stats_fused_kernel(const T* __restrict__ x, float* __restrict__ accum /*[2][C] persistent, zero*/,
                                      unsigned int* __restrict__ ticket, int64_t rows, int C, int cg, int rpb,
                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                      float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps,
                                      float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                      float* __restrict__ scale, float* __restrict__ shift, const float* __restrict__ pre_bias) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float sm[];               // [rpb][2][C]
    __shared__ bool is_last;
    const int sub = threadIdx.x % cg, r = threadIdx.x / cg;
    float K[8], a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        K[i] = running_mean ? running_mean[sub * 8 + i] - (pre_bias ? pre_bias[sub * 8 + i] : 0.f) : 0.f;
        a0[i] = 0.f; a1[i] = 0.f;
    }
    const int64_t stride = (int64_t)gridDim.x * rpb;
    int64_t row = (int64_t)blockIdx.x * rpb + r;
    constexpr int U = 4;                                          // 4 independent 16/32-byte loads in flight per thread
    for (; row + (U - 1) * stride < rows; row += U * stride) {
        Raw8<T> raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) ldraw(x + (row + u * stride) * C + sub * 8, raw[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float v[8];
            unpack8(raw[u], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float d = v[i] - K[i]; a0[i] += d; a1[i] += d * d; }
        }
    }
    for (; row < rows; row += stride) {
        float v[8];
        load8(x + row * C + sub * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[i] - K[i]; a0[i] += d; a1[i] += d * d; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sm[((size_t)r * 2 + 0) * C + sub * 8 + i] = a0[i];
        sm[((size_t)r * 2 + 1) * C + sub * 8 + i] = a1[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
        const int which = c / C, ch = c % C;
        float t = 0.f;
        for (int rr = 0; rr < rpb; ++rr) t += sm[((size_t)rr * 2 + which) * C + ch];
        atomicAdd(accum + which * C + ch, t);
    }
    if (!ticket) return;                // "raw" protocol: the consumer (bn_act_fwd_kernel, BnFin) finalises and clears the totals
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const float n = (float)rows;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float S = __ldcg(accum + c), Q = __ldcg(accum + C + c);
        const float md = S / n;
        const float m2 = fmaxf(Q - S * md, 0.f);
        const float k = running_mean ? running_mean[c] - (pre_bias ? pre_bias[c] : 0.f) : 0.f;
        bn_finalize_channel(n, k + md, m2, c, gamma, beta, running_mean, running_var, momentum, eps, mean_out, invstd_out, scale, shift,
                            pre_bias);
        accum[c] = 0.f;
        accum[C + c] = 0.f;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

// Chan-combine nparts partials -> stat[C][2] = (mean, M2), total[0] = count.  One WARP per channel.
__global__ void bn_combine_kernel(const float* __restrict__ part, const float* __restrict__ cnt, int nparts, int C,
                                  float* __restrict__ stat, float* __restrict__ total) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    float na, ma, qa;
    bn_warp_combine(part, cnt, nparts, C, c, lane, na, ma, qa);
    if (lane == 0) {
        stat[c * 2] = ma;
        stat[c * 2 + 1] = qa;
        if (c == 0) total[0] = na;
    }
}

__global__ void bn_finalize_kernel(const float* __restrict__ stat, const float* __restrict__ total,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float momentum, float eps, int C,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                   float* __restrict__ scale, float* __restrict__ shift, const float* __restrict__ pre_bias) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float n = total[0];
    const float mean = stat[c * 2];
    const float var = stat[c * 2 + 1] / n;                      // biased (normalisation)
    const float invstd = rsqrtf(var + eps);
    mean_out[c] = mean;
    invstd_out[c] = invstd;
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - mean * sc;
    if (running_mean) {
        const float unb = stat[c * 2 + 1] / fmaxf(n - 1.f, 1.f);  // unbiased (running estimate)
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (mean + (pre_bias ? pre_bias[c] : 0.f));
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
    }
}

__global__ void bn_eval_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps, int C,
                                      float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                      float* __restrict__ scale, float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float invstd = rsqrtf(rv[c] + eps);
    const float sc = gamma[c] * invstd;
    mean_out[c] = rm[c]; invstd_out[c] = invstd; scale[c] = sc; shift[c] = beta[c] - rm[c] * sc;
}

// ---------------------------------------------------------------------------------------------
// "Raw" protocol of the single-rank training path (accum != NULL): the statistics kernel only adds its per-block sums into the
// layer's persistent scratch and exits -- no fence, no ticket, no finalize block, i.e. three dependent global round trips less on
// every one of the 330 layers -- and THIS kernel turns the raw totals into scale/shift in its prologue (a handful of flops per
// thread).  The bookkeeping that must happen exactly once (saved mean/invstd/scale/shift for the backward pass, running
// statistics, clearing the scratch) is done by whichever block finishes LAST (ticket taken after the block's work, nobody waits).
struct BnFin {
    float* accum;                       // [2C] sum (x-K), sum (x-K)^2 ; NULL = legacy mode (scale/shift come in precomputed)
    unsigned int* ticket;
    const float* gamma; const float* beta;
    float* running_mean; float* running_var;      // may be NULL
    float momentum, eps;
    float* mean_out; float* invstd_out; float* scale_out; float* shift_out;
    const float* pre_bias;              // may be NULL
    float n;                            // rows
};

template <typename T, int ACT, bool RES>   // ACT: 0 none, 1 relu, 2 gelu
__global__ void bn_act_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                  const float* __restrict__ scale, const float* __restrict__ shift,
                                  int64_t rows, int C, int cg, int rpb, const BnFin fin) {
    pdl_wait();
    pdl_trigger();
    const int sub = threadIdx.x % cg, r = threadIdx.x / cg;
    float sc[8], sh[8];
    if (fin.accum) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = sub * 8 + i;
            const float K = fin.running_mean ? fin.running_mean[c] - (fin.pre_bias ? fin.pre_bias[c] : 0.f) : 0.f;
            // plain (L1-cached) loads on purpose: every thread of up to 1184 blocks reads the same 2C floats, and L1-bypassing
            // __ldcg turned that into a hot spot on ONE L2 slice (measured: 51 us instead of 10 for the whole kernel); the
            // totals were written by the previous kernel, so L1 cannot hold a stale copy
            const float S = fin.accum[c], Q = fin.accum[C + c];
            const float md = S / fin.n;
            const float m2 = fmaxf(Q - S * md, 0.f);
            const float invstd = rsqrtf(m2 / fin.n + fin.eps);
            sc[i] = fin.gamma[c] * invstd;
            sh[i] = fin.beta[c] - (K + md) * sc[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { sc[i] = scale[sub * 8 + i]; sh[i] = shift[sub * 8 + i]; }
    }
    const int64_t stride = (int64_t)gridDim.x * rpb;
    int64_t row = (int64_t)blockIdx.x * rpb + r;
    constexpr int U = 4;
    auto one = [&](const Raw8<T>& rx, const Raw8<T>& rr, int64_t off) {
        float v[8];
        unpack8(rx, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = v[i] * sc[i] + sh[i];
        if (RES) {
            float a[8];
            unpack8(rr, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += a[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (ACT == 1) v[i] = fmaxf(v[i], 0.f);
            if (ACT == 2) v[i] = gelu_t<T>(v[i]);
        }
        store8(y + off, v);
    };
    for (; row + (U - 1) * stride < rows; row += U * stride) {
        Raw8<T> rx[U], rr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t off = (row + u * stride) * C + sub * 8;
            ldraw(x + off, rx[u]);
            if (RES) ldraw(res + off, rr[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) one(rx[u], rr[u], (row + u * stride) * C + sub * 8);
    }
    for (; row < rows; row += stride) {
        const int64_t off = row * C + sub * 8;
        Raw8<T> rx, rr;
        ldraw(x + off, rx);
        if (RES) ldraw(res + off, rr);
        one(rx, rr, off);
    }
    if (!fin.accum) return;
    __shared__ bool is_last;
    __syncthreads();                     // every thread of the block has consumed its totals
    // (no fence in front of the ticket: the block's reads of the totals completed long ago -- their values fed the main loop --
    //  and nothing this block wrote is read by the last block; a __threadfence() here, behind the block's burst of stores, made
    //  the kernel 4-5x slower on the 1184-block grids: measured 51 vs 10 us)
    if (threadIdx.x == 0) is_last = (atomicAdd(fin.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();                     // all other blocks are past their reads of accum / running_mean
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float S = __ldcg(fin.accum + c), Q = __ldcg(fin.accum + C + c);
        const float md = S / fin.n;
        const float m2 = fmaxf(Q - S * md, 0.f);
        const float k = fin.running_mean ? fin.running_mean[c] - (fin.pre_bias ? fin.pre_bias[c] : 0.f) : 0.f;
        bn_finalize_channel(fin.n, k + md, m2, c, fin.gamma, fin.beta, fin.running_mean, fin.running_var, fin.momentum, fin.eps,
                            fin.mean_out, fin.invstd_out, fin.scale_out, fin.shift_out, fin.pre_bias);
        fin.accum[c] = 0.f;
        fin.accum[C + c] = 0.f;
    }
    if (threadIdx.x == 0) *fin.ticket = 0u;
}

// dz = dy * act'(z).  relu: mask from the saved output y (>0) when HAS_Y (residual layers), else recomputed from x; gelu: z
// recomputed from x.
template <typename T, int ACT, bool HAS_Y>
__device__ __forceinline__ void bn_dz(const Raw8<T>& rx, const Raw8<T>& ry, const Raw8<T>& rd,
                                      const float sc[8], const float sh[8], const float mu[8], const float is[8],
                                      float dz[8], float xh[8]) {
    float v[8];
    unpack8(rx, v);
    unpack8(rd, dz);
#pragma unroll
    for (int i = 0; i < 8; ++i) xh[i] = (v[i] - mu[i]) * is[i];
    if (ACT == 1) {
        if (HAS_Y) {
            float o[8];
            unpack8(ry, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) dz[i] = o[i] > 0.f ? dz[i] : 0.f;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) dz[i] = (v[i] * sc[i] + sh[i]) > 0.f ? dz[i] : 0.f;
        }
    }
    if (ACT == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dz[i] *= gelu_grad_t<T>(v[i] * sc[i] + sh[i]);
    }
}

template <typename T, int ACT, bool HAS_Y>
__global__ void bn_bwd_reduce_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ dy,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     float* __restrict__ sums /*[2][C]: sum dz, sum dz*xhat*/,
                                     float* __restrict__ accum /*[2][C] persistent zeroed scratch, or NULL: sums was zeroed by the caller*/,
                                     unsigned int* __restrict__ ticket, T* __restrict__ dz_out /*NULL or [rows][C]: dz kept for the apply pass*/,
                                     int64_t rows, int C, int cg, int rpb) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float sm[];           // [rpb][2][C]
    __shared__ bool is_last;
    const int sub = threadIdx.x % cg, r = threadIdx.x / cg;
    float sc[8], sh[8], mu[8], is[8], a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = scale[sub * 8 + i]; sh[i] = shift[sub * 8 + i]; mu[i] = mean[sub * 8 + i]; is[i] = invstd[sub * 8 + i];
        a0[i] = 0.f; a1[i] = 0.f;
    }
    const int64_t stride = (int64_t)gridDim.x * rpb;
    int64_t row = (int64_t)blockIdx.x * rpb + r;
    constexpr int U = 4;
    for (; row + (U - 1) * stride < rows; row += U * stride) {
        Raw8<T> rx[U], ry[U], rd[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t off = (row + u * stride) * C + sub * 8;
            ldraw(x + off, rx[u]);
            ldraw(dy + off, rd[u]);
            if (HAS_Y) ldraw(y + off, ry[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float dz[8], xh[8];
            bn_dz<T, ACT, HAS_Y>(rx[u], ry[u], rd[u], sc, sh, mu, is, dz, xh);
            if (dz_out) store8(dz_out + (row + u * stride) * C + sub * 8, dz);
#pragma unroll
            for (int i = 0; i < 8; ++i) { a0[i] += dz[i]; a1[i] += dz[i] * xh[i]; }
        }
    }
    for (; row < rows; row += stride) {
        const int64_t off = row * C + sub * 8;
        Raw8<T> rx, ry, rd;
        ldraw(x + off, rx);
        ldraw(dy + off, rd);
        if (HAS_Y) ldraw(y + off, ry);
        float dz[8], xh[8];
        bn_dz<T, ACT, HAS_Y>(rx, ry, rd, sc, sh, mu, is, dz, xh);
        if (dz_out) store8(dz_out + off, dz);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a0[i] += dz[i]; a1[i] += dz[i] * xh[i]; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sm[((size_t)r * 2 + 0) * C + sub * 8 + i] = a0[i];
        sm[((size_t)r * 2 + 1) * C + sub * 8 + i] = a1[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
        const int which = c / C, ch = c % C;
        float s = 0.f;
        for (int rr = 0; rr < rpb; ++rr) s += sm[((size_t)rr * 2 + which) * C + ch];
        atomicAdd((accum ? accum : sums) + which * C + ch, s);
    }
    if (!accum || !ticket) return;       // (accum without ticket: "raw" protocol, bn_bwd_apply_kernel publishes and clears the totals)
    // no memset node in front of the kernel (one launch + one dependency edge less per layer on a latency-bound chain): the
    // totals are built in a persistent scratch that every launch leaves zeroed; the last block to arrive publishes them
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
        sums[c] = __ldcg(accum + c);
        accum[c] = 0.f;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

template <typename T, int ACT, bool HAS_Y>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ dy,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ sums, float inv_count,
                                    T* __restrict__ dx, T* __restrict__ dres, int64_t rows, int C, int cg, int rpb,
                                    const float* __restrict__ local_sums, float* __restrict__ dgamma_acc, float* __restrict__ dbeta_acc,
                                    float* __restrict__ raw_accum /*NULL, or the scratch bn_bwd_reduce_kernel added into ("raw" protocol)*/,
                                    unsigned int* __restrict__ raw_ticket, float* __restrict__ sums_out /*NULL or [2C]*/) {
    pdl_wait();
    pdl_trigger();
    const int sub = threadIdx.x % cg, r = threadIdx.x / cg;
    if (!raw_accum && dgamma_acc && blockIdx.x == 0) {      // parameter gradients straight into the caller's (flat) grad buffer
        for (int c = threadIdx.x; c < C; c += blockDim.x) { dbeta_acc[c] += local_sums[c]; dgamma_acc[c] += local_sums[C + c]; }
    }
    float sc[8], sh[8], mu[8], is[8], m0[8], m1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = scale[sub * 8 + i]; sh[i] = shift[sub * 8 + i]; mu[i] = mean[sub * 8 + i]; is[i] = invstd[sub * 8 + i];
        if (raw_accum) {
            m0[i] = raw_accum[sub * 8 + i] * inv_count; m1[i] = raw_accum[C + sub * 8 + i] * inv_count;   // L1-cached on purpose, see BnFin
        } else {
            m0[i] = sums[sub * 8 + i] * inv_count; m1[i] = sums[C + sub * 8 + i] * inv_count;
        }
    }
    const int64_t stride = (int64_t)gridDim.x * rpb;
    int64_t row = (int64_t)blockIdx.x * rpb + r;
    constexpr int U = 2;
    auto one = [&](const Raw8<T>& rx, const Raw8<T>& ry, const Raw8<T>& rd, int64_t off) {
        float dz[8], xh[8], o[8];
        bn_dz<T, ACT, HAS_Y>(rx, ry, rd, sc, sh, mu, is, dz, xh);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = sc[i] * (dz[i] - m0[i] - xh[i] * m1[i]);
        store8(dx + off, o);
        if (dres) store8(dres + off, dz);
    };
    for (; row + (U - 1) * stride < rows; row += U * stride) {
        Raw8<T> rx[U], ry[U], rd[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t off = (row + u * stride) * C + sub * 8;
            ldraw(x + off, rx[u]);
            ldraw(dy + off, rd[u]);
            if (HAS_Y) ldraw(y + off, ry[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) one(rx[u], ry[u], rd[u], (row + u * stride) * C + sub * 8);
    }
    for (; row < rows; row += stride) {
        const int64_t off = row * C + sub * 8;
        Raw8<T> rx, ry, rd;
        ldraw(x + off, rx);
        ldraw(dy + off, rd);
        if (HAS_Y) ldraw(y + off, ry);
        one(rx, ry, rd, off);
    }
    if (!raw_accum) return;
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(raw_ticket, 1u) == gridDim.x - 1);      // no fence needed, see bn_act_fwd_kernel
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {      // once per launch: publish, accumulate dgamma/dbeta, clear the scratch
        const float S0 = __ldcg(raw_accum + c), S1 = __ldcg(raw_accum + C + c);
        if (sums_out) { sums_out[c] = S0; sums_out[C + c] = S1; }
        if (dgamma_acc) { dbeta_acc[c] += S0; dgamma_acc[c] += S1; }
        raw_accum[c] = 0.f;
        raw_accum[C + c] = 0.f;
    }
    if (threadIdx.x == 0) *raw_ticket = 0u;
}

// apply pass when the reduce pass kept dz = dy*act'(z) (GELU layers: the derivative costs ~30 instructions and two MUFU ops per
// element, which made BOTH passes issue-bound at ~3x their HBM time; with dz stored once the second pass is a plain stream)
template <typename T>
__global__ void bn_bwd_apply_dz_kernel(const T* __restrict__ x, const T* __restrict__ dz, const float* __restrict__ scale,
                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                       const float* __restrict__ sums, float inv_count, T* __restrict__ dx,
                                       int64_t rows, int C, int cg, int rpb,
                                       const float* __restrict__ local_sums, float* __restrict__ dgamma_acc, float* __restrict__ dbeta_acc) {
    pdl_wait();
    pdl_trigger();
    const int sub = threadIdx.x % cg, r = threadIdx.x / cg;
    if (dgamma_acc && blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) { dbeta_acc[c] += local_sums[c]; dgamma_acc[c] += local_sums[C + c]; }
    }
    float sc[8], mu[8], is[8], m0[8], m1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = scale[sub * 8 + i]; mu[i] = mean[sub * 8 + i]; is[i] = invstd[sub * 8 + i];
        m0[i] = sums[sub * 8 + i] * inv_count; m1[i] = sums[C + sub * 8 + i] * inv_count;
    }
    const int64_t stride = (int64_t)gridDim.x * rpb;
    int64_t row = (int64_t)blockIdx.x * rpb + r;
    constexpr int U = 4;
    auto one = [&](const Raw8<T>& rx, const Raw8<T>& rz, int64_t off) {
        float v[8], d[8], o[8];
        unpack8(rx, v);
        unpack8(rz, d);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = sc[i] * (d[i] - m0[i] - (v[i] - mu[i]) * is[i] * m1[i]);
        store8(dx + off, o);
    };
    for (; row + (U - 1) * stride < rows; row += U * stride) {
        Raw8<T> rx[U], rz[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t off = (row + u * stride) * C + sub * 8;
            ldraw(x + off, rx[u]);
            ldraw(dz + off, rz[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) one(rx[u], rz[u], (row + u * stride) * C + sub * 8);
    }
    for (; row < rows; row += stride) {
        const int64_t off = row * C + sub * 8;
        Raw8<T> rx, rz;
        ldraw(x + off, rx);
        ldraw(dz + off, rz);
        one(rx, rz, off);
    }
}

// (A one-launch BatchNorm -- statistics -> device-wide spin barrier -> apply, and reduce -> barrier -> apply -- lived here through
// round 2: correct and tested, but the barrier cost more than the launch it saved (13-32 us per kernel vs ~10 + ~5 us for the two
// split kernels, 382 vs 418 img/s on the B=16 step), so it was removed; see the history at "one-launch BN kernels".)

}  // namespace rss

using namespace rss;

extern "C" int rss_bn_stats_nparts(int64_t rows, int C) {
    if (C <= 0 || C % 8) return RSS_ERR_SHAPE;
    const BnGeom g = bn_geom(C);
    return bn_grid(rows, g.rpb * 8, 2);
}

extern "C" int rss_bn_stats(const void* x, float* partials, float* counts, int64_t rows, int C, int dtype, cudaStream_t st) {
    if (C <= 0 || C % 8 || C > 2048 || rows <= 0) return RSS_ERR_SHAPE;
    const BnGeom g = bn_geom(C);
    const int grid = rss_bn_stats_nparts(rows, C);
    const size_t smem = (size_t)g.rpb * C * 3 * sizeof(float);
    RSS_DISPATCH_DTYPE(dtype, bn_stats_kernel<T><<<grid, g.threads, smem, st>>>((const T*)x, partials, counts, rows, C, g.cg, g.rpb));
    return check_launch();
}

extern "C" int rss_bn_stats_fused(const void* x, float* accum_scratch, unsigned int* ticket, int64_t rows, int C, int dtype,
                                  const float* gamma, const float* beta, float* running_mean, float* running_var,
                                  float momentum, float eps, float* mean_out, float* invstd_out, float* scale, float* shift,
                                  const float* pre_bias, cudaStream_t st) {
    if (C <= 0 || C % 8 || C > 2048 || rows <= 0 || !ticket || !accum_scratch) return RSS_ERR_SHAPE;
    const BnGeom g = bn_geom(C);
    const int grid = bn_grid(rows, g.rpb * 8, bn_ticket_bpsm_for(rows, C, dtype));
    const size_t smem = (size_t)g.rpb * C * 2 * sizeof(float);
    RSS_DISPATCH_DTYPE(dtype, launch_k(bn_stats_fused_kernel<T>, grid, g.threads, smem, st, (const T*)x, accum_scratch, ticket, rows, C, g.cg, g.rpb,
                       gamma, beta, running_mean, running_var, momentum, eps, mean_out, invstd_out, scale, shift, pre_bias));
    return check_launch();
}

extern "C" int rss_bn_combine(const float* partials, const float* counts, int nparts, int C, float* stat, float* total,
                              cudaStream_t st) {
    if (C <= 0 || nparts <= 0) return RSS_ERR_SHAPE;
    bn_combine_kernel<<<(C * 32 + 255) / 256, 256, 0, st>>>(partials, counts, nparts, C, stat, total);
    return check_launch();
}

extern "C" int rss_bn_finalize(const float* stat, const float* total, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float momentum, float eps, int C,
                               float* mean_out, float* invstd_out, float* scale, float* shift, const float* pre_bias, cudaStream_t st) {
    if (C <= 0) return RSS_ERR_SHAPE;
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(stat, total, gamma, beta, running_mean, running_var, momentum, eps, C,
                                                        mean_out, invstd_out, scale, shift, pre_bias);
    return check_launch();
}

extern "C" int rss_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                                  float eps, int C, float* mean_out, float* invstd_out, float* scale, float* shift, cudaStream_t st) {
    if (C <= 0) return RSS_ERR_SHAPE;
    bn_eval_affine_kernel<<<(C + 127) / 128, 128, 0, st>>>(gamma, beta, running_mean, running_var, eps, C, mean_out, invstd_out, scale, shift);
    return check_launch();
}

#define BN_ACT_SWITCH(KERNEL, FLAG, ...)                                             \
    switch (act) {                                                                   \
        case RSS_ACT_NONE: if (FLAG) KERNEL<T, 0, true> __VA_ARGS__; else KERNEL<T, 0, false> __VA_ARGS__; break; \
        case RSS_ACT_RELU: if (FLAG) KERNEL<T, 1, true> __VA_ARGS__; else KERNEL<T, 1, false> __VA_ARGS__; break; \
        case RSS_ACT_GELU: KERNEL<T, 2, false> __VA_ARGS__; break;                   \
        default: return RSS_ERR_SHAPE;                                               \
    }
// same dispatch through launch_k (programmatic dependent launch, common.cuh)
#define BN_ACT_LAUNCH(KERNEL, FLAG, GRID, BLOCK, SMEM, ST, ...)                      \
    switch (act) {                                                                   \
        case RSS_ACT_NONE: if (FLAG) launch_k(KERNEL<T, 0, true>, GRID, BLOCK, SMEM, ST, __VA_ARGS__); else launch_k(KERNEL<T, 0, false>, GRID, BLOCK, SMEM, ST, __VA_ARGS__); break; \
        case RSS_ACT_RELU: if (FLAG) launch_k(KERNEL<T, 1, true>, GRID, BLOCK, SMEM, ST, __VA_ARGS__); else launch_k(KERNEL<T, 1, false>, GRID, BLOCK, SMEM, ST, __VA_ARGS__); break; \
        case RSS_ACT_GELU: launch_k(KERNEL<T, 2, false>, GRID, BLOCK, SMEM, ST, __VA_ARGS__); break; \
        default: return RSS_ERR_SHAPE;                                               \
    }

extern "C" int rss_bn_act_fwd(const void* x, const void* residual, void* y, const float* scale, const float* shift,
                              int64_t rows, int C, int act, int dtype, cudaStream_t st) {
    if (C <= 0 || C % 8 || rows <= 0) return RSS_ERR_SHAPE;
    if (residual && act == RSS_ACT_GELU) return RSS_ERR_SHAPE;   // not a pattern of the reference (backward would need the residual)
    const BnGeom g = bn_geom(C);
    const int grid = bn_grid(rows, g.rpb * 4, bn_apply_bpsm_for(rows, C, dtype));
    RSS_DISPATCH_DTYPE(dtype, BN_ACT_LAUNCH(bn_act_fwd_kernel, residual != nullptr, grid, g.threads, 0, st, (const T*)x, (const T*)residual, (T*)y, scale, shift, rows, C, g.cg, g.rpb, BnFin{}));
    return check_launch();
}

// ---- "raw" protocol (single-rank training path): statistics kernel = sums only; the apply kernel finalises (see BnFin) ----
extern "C" int rss_bn_stats_raw(const void* x, float* accum_scratch, int64_t rows, int C, int dtype,
                                const float* running_mean, const float* pre_bias, cudaStream_t st) {
    if (C <= 0 || C % 8 || C > 2048 || rows <= 0 || !accum_scratch) return RSS_ERR_SHAPE;
    const BnGeom g = bn_geom(C);
    const int grid = bn_grid(rows, g.rpb * 8, bn_ticket_bpsm());
    const size_t smem = (size_t)g.rpb * C * 2 * sizeof(float);
    RSS_DISPATCH_DTYPE(dtype, bn_stats_fused_kernel<T><<<grid, g.threads, smem, st>>>((const T*)x, accum_scratch, nullptr, rows, C, g.cg, g.rpb,
                       nullptr, nullptr, const_cast<float*>(running_mean), nullptr, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr, pre_bias));
    return check_launch();
}

extern "C" int rss_bn_act_fwd_raw(const void* x, const void* residual, void* y, float* accum_scratch, unsigned int* ticket,
                                  int64_t rows, int C, int act, int dtype, const float* gamma, const float* beta,
                                  float* running_mean, float* running_var, float momentum, float eps,
                                  float* mean_out, float* invstd_out, float* scale, float* shift, const float* pre_bias,
                                  cudaStream_t st) {
    if (C <= 0 || C % 8 || rows <= 0 || !accum_scratch || !ticket) return RSS_ERR_SHAPE;
    if (residual && act == RSS_ACT_GELU) return RSS_ERR_SHAPE;
    const BnGeom g = bn_geom(C);
    const int grid = bn_grid(rows, g.rpb * 4, bn_apply_bpsm_for(rows, C, dtype));
    BnFin fin;
    fin.accum = accum_scratch; fin.ticket = ticket; fin.gamma = gamma; fin.beta = beta; fin.running_mean = running_mean;
    fin.running_var = running_var; fin.momentum = momentum; fin.eps = eps; fin.mean_out = mean_out; fin.invstd_out = invstd_out;
    fin.scale_out = scale; fin.shift_out = shift; fin.pre_bias = pre_bias; fin.n = (float)rows;
    RSS_DISPATCH_DTYPE(dtype, BN_ACT_SWITCH(bn_act_fwd_kernel, residual != nullptr, <<<grid, g.threads, 0, st>>>((const T*)x, (const T*)residual, (T*)y, nullptr, nullptr, rows, C, g.cg, g.rpb, fin)));
    return check_launch();
}

// accum_scratch/ticket: optional persistent per-layer scratch (float[2C] + one counter, zero on entry, left zero): with it the
// launch needs no memset of `sums` in front of it.  Both NULL: `sums` is cleared by a memset node first.
extern "C" int rss_bn_bwd_reduce_ws(const void* x, const void* y, const void* dy, const float* scale, const float* shift,
                                    const float* mean, const float* invstd, float* sums, float* accum_scratch, unsigned int* ticket,
                                    void* dz_out, int64_t rows, int C, int act, int dtype, cudaStream_t st) {
    if (C <= 0 || C % 8 || rows <= 0 || (ticket && !accum_scratch)) return RSS_ERR_SHAPE;
    const BnGeom g = bn_geom(C);
    const int grid = bn_grid(rows, g.rpb * 8, bn_ticket_bpsm_for(rows, C, dtype));
    const size_t smem = (size_t)g.rpb * 2 * C * sizeof(float);
    if (!accum_scratch) {
        cudaError_t e = cudaMemsetAsync(sums, 0, 2 * C * sizeof(float), st);
        if (e != cudaSuccess) { g_last_cuda_error = (int)e; return RSS_ERR_CUDA; }
    }
    RSS_DISPATCH_DTYPE(dtype, BN_ACT_LAUNCH(bn_bwd_reduce_kernel, y != nullptr && act == RSS_ACT_RELU, grid, g.threads, smem, st, (const T*)x, (const T*)y, (const T*)dy, scale, shift, mean, invstd, sums, accum_scratch, ticket, (T*)dz_out, rows, C, g.cg, g.rpb));
    return check_launch();
}

extern "C" int rss_bn_bwd_reduce(const void* x, const void* y, const void* dy, const float* scale, const float* shift,
                                 const float* mean, const float* invstd, float* sums, int64_t rows, int C, int act, int dtype,
                                 cudaStream_t st) {
    return rss_bn_bwd_reduce_ws(x, y, dy, scale, shift, mean, invstd, sums, nullptr, nullptr, nullptr, rows, C, act, dtype, st);
}

extern "C" int rss_bn_bwd_apply(const void* x, const void* y, const void* dy, const float* scale, const float* shift,
                                const float* mean, const float* invstd, const float* sums, float inv_count,
                                void* dx, void* dres, int64_t rows, int C, int act, int dtype,
                                const float* local_sums, float* dgamma_acc, float* dbeta_acc, cudaStream_t st) {
    if (C <= 0 || C % 8 || rows <= 0) return RSS_ERR_SHAPE;
    if (act == RSS_ACT_RELU && dres && !y) return RSS_ERR_SHAPE;     // residual layers must pass the saved output
    const BnGeom g = bn_geom(C);
    const int grid = bn_grid(rows, g.rpb * 4, bn_apply_bpsm_for(rows, C, dtype));
    RSS_DISPATCH_DTYPE(dtype, BN_ACT_LAUNCH(bn_bwd_apply_kernel, y != nullptr && act == RSS_ACT_RELU, grid, g.threads, 0, st, (const T*)x, (const T*)y, (const T*)dy, scale, shift, mean, invstd, sums, inv_count, (T*)dx, (T*)dres, rows, C, g.cg, g.rpb,
                                            local_sums, dgamma_acc, dbeta_acc, (float*)nullptr, (unsigned int*)nullptr, (float*)nullptr));
    return check_launch();
}

// "raw" protocol backward: rss_bn_bwd_reduce_ws(accum_scratch, ticket = NULL) only adds the per-block sums into the scratch;
// this apply pass reads them there, and its last block publishes sums_out (optional), accumulates dgamma/dbeta and clears the scratch
extern "C" int rss_bn_bwd_apply_raw(const void* x, const void* y, const void* dy, const float* scale, const float* shift,
                                    const float* mean, const float* invstd, float* accum_scratch, unsigned int* ticket,
                                    float inv_count, void* dx, void* dres, int64_t rows, int C, int act, int dtype,
                                    float* sums_out, float* dgamma_acc, float* dbeta_acc, cudaStream_t st) {
    if (C <= 0 || C % 8 || rows <= 0 || !accum_scratch || !ticket) return RSS_ERR_SHAPE;
    if (act == RSS_ACT_RELU && dres && !y) return RSS_ERR_SHAPE;
    const BnGeom g = bn_geom(C);
    const int grid = bn_grid(rows, g.rpb * 4, bn_apply_bpsm_for(rows, C, dtype));
    RSS_DISPATCH_DTYPE(dtype, BN_ACT_SWITCH(bn_bwd_apply_kernel, y != nullptr && act == RSS_ACT_RELU, <<<grid, g.threads, 0, st>>>((const T*)x, (const T*)y, (const T*)dy, scale, shift, mean, invstd, nullptr, inv_count, (T*)dx, (T*)dres, rows, C, g.cg, g.rpb,
                                                                                                           nullptr, dgamma_acc, dbeta_acc, accum_scratch, ticket, sums_out)));
    return check_launch();
}

// dx = scale*(dz - sum_dz/n - xhat*sum_dzxhat/n) from the dz = dy*act'(.) kept by rss_bn_bwd_reduce_ws(dz_out)
extern "C" int rss_bn_bwd_apply_dz(const void* x, const void* dz, const float* scale, const float* mean, const float* invstd,
                                   const float* sums, float inv_count, void* dx, int64_t rows, int C, int dtype,
                                   const float* local_sums, float* dgamma_acc, float* dbeta_acc, cudaStream_t st) {
    if (C <= 0 || C % 8 || rows <= 0 || !dz) return RSS_ERR_SHAPE;
    const BnGeom g = bn_geom(C);
    const int grid = bn_grid(rows, g.rpb * 4, bn_apply_bpsm_for(rows, C, dtype));
    RSS_DISPATCH_DTYPE(dtype, launch_k(bn_bwd_apply_dz_kernel<T>, grid, g.threads, 0, st, (const T*)x, (const T*)dz, scale, mean, invstd, sums,
                       inv_count, (T*)dx, rows, C, g.cg, g.rpb, local_sums, dgamma_acc, dbeta_acc));
    return check_launch();
}
