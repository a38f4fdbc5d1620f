// BatchNorm(+ReLU)(+residual) in ONE launch for the small (L2-resident) activations of the HRNet branches 1-3, fuse and transition
// layers (_hrnet_rssformer.py:216-287,361-405,512-546): statistics pass + apply pass (forward) / reduce pass + apply pass
// (backward) without a device-wide synchronisation.
//
// Why: these ~200 layers are 2-8 MB tensors that live in the 126 MB L2; the two-kernel protocol of bn.cu costs 12.5 + 7 us
// (forward) and 13 + 13 us (backward) on them at 8-16 % of DRAM throughput (profiles/ncu_r2_full_bn_convcf.csv): launch latency,
// global atomics, a device-wide ticket and a last-block finalise per pass.  A one-launch version with a device-wide spin barrier
// was measured slower in round 1 (all blocks must be co-resident and spin while other streams want the SMs).
//
// Here the work is split by CHANNEL, not by row: a thread-block CLUSTER of 8 CTAs owns a slice of 16 channels (32 bytes of bf16 per
// pixel: whole sectors) over ALL rows, so a channel's reduction never leaves the cluster: per-CTA partials are exchanged through
// distributed shared memory (cluster.map_shared_rank) around two hardware cluster barriers -- no atomics, no tickets, no scratch,
// deterministic summation order.  The second pass re-reads the slice from L2.  Grid = 8 x C/16 CTAs (32 for C = 64 ... 128 for
// C = 256): small enough to slot in between the kernels of the other streams.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace rss {

constexpr int kClSize = 8;            // CTAs per cluster (portable maximum)
constexpr int kClThreads = 512;      // x 4 rows in flight per thread: the passes are L2-latency bound (32-128 CTAs), not bandwidth bound
constexpr int kClCh = 16;             // channels per cluster

template <typename T> struct Row16;   // 16 consecutive channels of one row
template <> struct Row16<__nv_bfloat16> {
    uint4 a, b;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { a = __ldg(reinterpret_cast<const uint4*>(p)); b = __ldg(reinterpret_cast<const uint4*>(p) + 1); }
    __device__ __forceinline__ void unpack(float v[16]) const {
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float v[16]) { store8(p, v); store8(p + 8, v + 8); }
};
template <> struct Row16<float> {
    float4 q[4];
    __device__ __forceinline__ void load(const float* p) {
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = __ldg(reinterpret_cast<const float4*>(p) + i);
    }
    __device__ __forceinline__ void unpack(float v[16]) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[4 * i] = q[i].x; v[4 * i + 1] = q[i].y; v[4 * i + 2] = q[i].z; v[4 * i + 3] = q[i].w; }
    }
    static __device__ __forceinline__ void store(float* p, const float v[16]) { store8(p, v); store8(p + 8, v + 8); }
};

// block-level sum of 32 per-thread values -> part[32] (shared), then the cluster-wide totals in tot[32] of every CTA
__device__ __forceinline__ void cluster_totals(float vals[32], float (*wpart)[32], float* part, float* tot, cg::cluster_group& cl) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float s = warp_sum(vals[i]);
        if (lane == 0) wpart[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kClThreads / 32; ++w) t += wpart[w][threadIdx.x];
        part[threadIdx.x] = t;
    }
    cl.sync();                                            // every CTA's partial is visible cluster-wide
    if (threadIdx.x < 32) {
        float t = 0.f;
        for (int r = 0; r < kClSize; ++r) t += cl.map_shared_rank(part, r)[threadIdx.x];      // fixed order: deterministic
        tot[threadIdx.x] = t;
    }
    cl.sync();                                            // nobody leaves (or overwrites `part`) while a peer may still read it
}

template <typename T, int ACT, bool RES>
__global__ void __cluster_dims__(kClSize, 1, 1) __launch_bounds__(kClThreads)
bn_cluster_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y, int64_t rows, int C,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
                      float* __restrict__ running_var, float momentum, float eps, float* __restrict__ mean_out,
                      float* __restrict__ invstd_out, float* __restrict__ scale_out, float* __restrict__ shift_out,
                      const float* __restrict__ pre_bias) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ float wpart[kClThreads / 32][32], part[32], tot[32], aff[2][kClCh];
    const int rank = (int)cl.block_rank(), c0 = (blockIdx.x / kClSize) * kClCh;
    const int64_t rpc = (rows + kClSize - 1) / kClSize;
    const int64_t r_lo = (int64_t)rank * rpc, r_hi = r_lo + rpc < rows ? r_lo + rpc : rows;
    float K[kClCh], vals[32];
#pragma unroll
    for (int i = 0; i < kClCh; ++i) K[i] = running_mean ? running_mean[c0 + i] - (pre_bias ? pre_bias[c0 + i] : 0.f) : 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) vals[i] = 0.f;
    int64_t row = r_lo + threadIdx.x;
    constexpr int U = 4;
    for (; row + (U - 1) * kClThreads < r_hi; row += U * kClThreads) {
        Row16<T> rr_[U];
#pragma unroll
        for (int u = 0; u < U; ++u) rr_[u].load(x + (row + u * kClThreads) * C + c0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float v[16];
            rr_[u].unpack(v);
#pragma unroll
            for (int i = 0; i < kClCh; ++i) { const float d = v[i] - K[i]; vals[i] += d; vals[kClCh + i] = fmaf(d, d, vals[kClCh + i]); }
        }
    }
    for (; row < r_hi; row += kClThreads) {
        Row16<T> ra;
        ra.load(x + row * C + c0);
        float v[16];
        ra.unpack(v);
#pragma unroll
        for (int i = 0; i < kClCh; ++i) { const float d = v[i] - K[i]; vals[i] += d; vals[kClCh + i] = fmaf(d, d, vals[kClCh + i]); }
    }
    cluster_totals(vals, wpart, part, tot, cl);
    if (threadIdx.x < kClCh) {
        const int c = c0 + threadIdx.x;
        const float n = (float)rows, S = tot[threadIdx.x], Q = tot[kClCh + threadIdx.x];
        const float md = S / n, m2 = fmaxf(Q - S * md, 0.f);
        const float k = running_mean ? running_mean[c] - (pre_bias ? pre_bias[c] : 0.f) : 0.f;
        const float mean = k + md, invstd = rsqrtf(m2 / n + eps);
        const float sc = gamma[c] * invstd;
        aff[0][threadIdx.x] = sc;
        aff[1][threadIdx.x] = beta[c] - mean * sc;
        if (rank == 0) {                                   // one writer per channel
            mean_out[c] = mean; invstd_out[c] = invstd; scale_out[c] = sc; shift_out[c] = beta[c] - mean * sc;
            if (running_mean) {
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (mean + (pre_bias ? pre_bias[c] : 0.f));
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (m2 / fmaxf(n - 1.f, 1.f));
            }
        }
    }
    __syncthreads();
    float sc[kClCh], sh[kClCh];
#pragma unroll
    for (int i = 0; i < kClCh; ++i) { sc[i] = aff[0][i]; sh[i] = aff[1][i]; }
    auto apply = [&](const Row16<T>& ra, const Row16<T>& rr, int64_t off) {
        float v[16];
        ra.unpack(v);
#pragma unroll
        for (int i = 0; i < kClCh; ++i) v[i] = fmaf(v[i], sc[i], sh[i]);
        if (RES) {
            float a[16];
            rr.unpack(a);
#pragma unroll
            for (int i = 0; i < kClCh; ++i) v[i] += a[i];
        }
        if (ACT == 1) {
#pragma unroll
            for (int i = 0; i < kClCh; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        Row16<T>::store(y + off, v);
    };
    row = r_lo + threadIdx.x;                                                // second pass: the slice comes back from L2
    for (; row + (U - 1) * kClThreads < r_hi; row += U * kClThreads) {
        Row16<T> ra[U], rr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            ra[u].load(x + (row + u * kClThreads) * C + c0);
            if (RES) rr[u].load(res + (row + u * kClThreads) * C + c0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) apply(ra[u], rr[u], (row + u * kClThreads) * C + c0);
    }
    for (; row < r_hi; row += kClThreads) {
        Row16<T> ra, rr;
        ra.load(x + row * C + c0);
        if (RES) rr.load(res + row * C + c0);
        apply(ra, rr, row * C + c0);
    }
}

template <typename T, bool HAS_Y> struct ClRows {
    Row16<T> rx, rd, ry;
    __device__ __forceinline__ void load(const T* x, const T* y, const T* dy, int64_t off) {
        rx.load(x + off);
        rd.load(dy + off);
        if (HAS_Y) ry.load(y + off);
    }
};

template <typename T, int ACT, bool HAS_Y>
__device__ __forceinline__ void cl_dz(const ClRows<T, HAS_Y>& in, const float sc[16], const float sh[16],
                                      const float mu[16], const float is[16], float dz[16], float xh[16]) {
    const Row16<T>& rx = in.rx; const Row16<T>& rd = in.rd; const Row16<T>& ry = in.ry;
    float v[16];
    rx.unpack(v);
    rd.unpack(dz);
#pragma unroll
    for (int i = 0; i < 16; ++i) xh[i] = (v[i] - mu[i]) * is[i];
    if (ACT == 1) {
        if (HAS_Y) {
            float o[16];
            ry.unpack(o);
#pragma unroll
            for (int i = 0; i < 16; ++i) dz[i] = o[i] > 0.f ? dz[i] : 0.f;
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) dz[i] = fmaf(v[i], sc[i], sh[i]) > 0.f ? dz[i] : 0.f;
        }
    }
}

template <typename T, int ACT, bool HAS_Y>
__global__ void __cluster_dims__(kClSize, 1, 1) __launch_bounds__(kClThreads)
bn_cluster_bwd_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ dy, const float* __restrict__ scale,
                      const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                      T* __restrict__ dx, T* __restrict__ dres, int64_t rows, int C, float* __restrict__ sums_out,
                      float* __restrict__ dgamma_acc, float* __restrict__ dbeta_acc) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ float wpart[kClThreads / 32][32], part[32], tot[32];
    const int rank = (int)cl.block_rank(), c0 = (blockIdx.x / kClSize) * kClCh;
    const int64_t rpc = (rows + kClSize - 1) / kClSize;
    const int64_t r_lo = (int64_t)rank * rpc, r_hi = r_lo + rpc < rows ? r_lo + rpc : rows;
    float sc[kClCh], sh[kClCh], mu[kClCh], is[kClCh], vals[32];
#pragma unroll
    for (int i = 0; i < kClCh; ++i) { sc[i] = scale[c0 + i]; sh[i] = shift[c0 + i]; mu[i] = mean[c0 + i]; is[i] = invstd[c0 + i]; }
#pragma unroll
    for (int i = 0; i < 32; ++i) vals[i] = 0.f;
    constexpr int U = 2;                                                       // 2 rows x (x, dy[, y]) in flight per thread
    int64_t row = r_lo + threadIdx.x;
    for (; row + (U - 1) * kClThreads < r_hi; row += U * kClThreads) {
        ClRows<T, HAS_Y> in[U];
#pragma unroll
        for (int u = 0; u < U; ++u) in[u].load(x, y, dy, (row + u * kClThreads) * C + c0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float dz[16], xh[16];
            cl_dz<T, ACT, HAS_Y>(in[u], sc, sh, mu, is, dz, xh);
#pragma unroll
            for (int i = 0; i < kClCh; ++i) { vals[i] += dz[i]; vals[kClCh + i] = fmaf(dz[i], xh[i], vals[kClCh + i]); }
        }
    }
    for (; row < r_hi; row += kClThreads) {
        ClRows<T, HAS_Y> in;
        in.load(x, y, dy, row * C + c0);
        float dz[16], xh[16];
        cl_dz<T, ACT, HAS_Y>(in, sc, sh, mu, is, dz, xh);
#pragma unroll
        for (int i = 0; i < kClCh; ++i) { vals[i] += dz[i]; vals[kClCh + i] = fmaf(dz[i], xh[i], vals[kClCh + i]); }
    }
    cluster_totals(vals, wpart, part, tot, cl);
    if (rank == 0 && threadIdx.x < kClCh) {                // one writer per channel: plain read-modify-write
        const int c = c0 + threadIdx.x;
        if (sums_out) { sums_out[c] = tot[threadIdx.x]; sums_out[C + c] = tot[kClCh + threadIdx.x]; }
        if (dgamma_acc) { dbeta_acc[c] += tot[threadIdx.x]; dgamma_acc[c] += tot[kClCh + threadIdx.x]; }
    }
    const float inv_n = 1.f / (float)rows;
    float m0[kClCh], m1[kClCh];
#pragma unroll
    for (int i = 0; i < kClCh; ++i) { m0[i] = tot[i] * inv_n; m1[i] = tot[kClCh + i] * inv_n; }
    auto apply = [&](const ClRows<T, HAS_Y>& in, int64_t off) {
        float dz[16], xh[16], o[16];
        cl_dz<T, ACT, HAS_Y>(in, sc, sh, mu, is, dz, xh);
#pragma unroll
        for (int i = 0; i < kClCh; ++i) o[i] = sc[i] * (dz[i] - m0[i] - xh[i] * m1[i]);
        Row16<T>::store(dx + off, o);
        if (dres) Row16<T>::store(dres + off, dz);
    };
    row = r_lo + threadIdx.x;
    for (; row + (U - 1) * kClThreads < r_hi; row += U * kClThreads) {
        ClRows<T, HAS_Y> in[U];
#pragma unroll
        for (int u = 0; u < U; ++u) in[u].load(x, y, dy, (row + u * kClThreads) * C + c0);
#pragma unroll
        for (int u = 0; u < U; ++u) apply(in[u], (row + u * kClThreads) * C + c0);
    }
    for (; row < r_hi; row += kClThreads) {
        ClRows<T, HAS_Y> in;
        in.load(x, y, dy, row * C + c0);
        apply(in, row * C + c0);
    }
}

}  // namespace rss

using namespace rss;

// 1 when the one-launch cluster kernels take this layer: whole 16-channel slices, no GELU (the GELU layers are the 17-67 MB FFN
// tensors), and small enough that 8 x C/16 CTAs re-reading their slice from L2 beat the row-parallel two-kernel protocol
extern "C" int rss_bn_cluster_supported(int64_t rows, int C, int act, int dtype) {
    if (rows <= 0 || C < 64 || C % kClCh || (act != RSS_ACT_NONE && act != RSS_ACT_RELU)) return 0;
    const int64_t bytes = rows * C * (dtype == RSS_F32 ? 4 : 2);
    return bytes <= (int64_t)12 << 20;
}

#define CL_SWITCH(KERNEL, FLAG, ...)                                                                        \
    if (act == RSS_ACT_NONE) { if (FLAG) KERNEL<T, 0, true> __VA_ARGS__; else KERNEL<T, 0, false> __VA_ARGS__; } \
    else { if (FLAG) KERNEL<T, 1, true> __VA_ARGS__; else KERNEL<T, 1, false> __VA_ARGS__; }

extern "C" int rss_bn_cluster_fwd(const void* x, const void* residual, void* y, int64_t rows, int C, int act, int dtype,
                                  const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum,
                                  float eps, float* mean_out, float* invstd_out, float* scale, float* shift, const float* pre_bias,
                                  cudaStream_t st) {
    if (!rss_bn_cluster_supported(rows, C, act, dtype)) return RSS_ERR_SHAPE;
    const int grid = kClSize * (C / kClCh);
    RSS_DISPATCH_DTYPE(dtype, CL_SWITCH(bn_cluster_fwd_kernel, residual != nullptr, <<<grid, kClThreads, 0, st>>>(
        (const T*)x, (const T*)residual, (T*)y, rows, C, gamma, beta, running_mean, running_var, momentum, eps, mean_out, invstd_out,
        scale, shift, pre_bias)));
    return check_launch();
}

extern "C" int rss_bn_cluster_bwd(const void* x, const void* y, const void* dy, const float* scale, const float* shift,
                                  const float* mean, const float* invstd, void* dx, void* dres, int64_t rows, int C, int act,
                                  int dtype, float* sums_out, float* dgamma_acc, float* dbeta_acc, cudaStream_t st) {
    if (!rss_bn_cluster_supported(rows, C, act, dtype)) return RSS_ERR_SHAPE;
    if (act == RSS_ACT_RELU && dres && !y) return RSS_ERR_SHAPE;     // residual layers must pass the saved output
    const int grid = kClSize * (C / kClCh);
    RSS_DISPATCH_DTYPE(dtype, CL_SWITCH(bn_cluster_bwd_kernel, y != nullptr && act == RSS_ACT_RELU, <<<grid, kClThreads, 0, st>>>(
        (const T*)x, (const T*)y, (const T*)dy, scale, shift, mean, invstd, (T*)dx, (T*)dres, rows, C, sums_out, dgamma_acc, dbeta_acc)));
    return check_launch();
}
