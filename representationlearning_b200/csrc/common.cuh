// Shared device helpers for the RSSFormer sm_100a kernels.
// Layout convention everywhere: activations are NHWC (== token-major (B, H*W, C)), dtype T in
// {float, __nv_bfloat16}; all arithmetic and all reductions are fp32; parameters are fp32.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>
#include "../../include/rss_b200.h"

namespace rss {

extern int g_last_cuda_error;

inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return RSS_ERR_CUDA; }
    return RSS_OK;
}

#define RSS_DISPATCH_DTYPE(dtype, ...)                                   \
    do {                                                                 \
        if ((dtype) == RSS_F32) { using T = float; __VA_ARGS__; }        \
        else if ((dtype) == RSS_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
        else return RSS_ERR_DTYPE;                                       \
    } while (0)

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive elements (16 B of bf16 / 32 B of fp32); p must be 16-byte aligned.
__device__ __forceinline__ void load8(const float* __restrict__ p, float v[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* __restrict__ p, float v[8]) {
    uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
// raw (not yet converted) 8-element vector: lets a thread issue several loads before touching the data
template <typename T> struct Raw8;
template <> struct Raw8<float> { float4 a, b; };
template <> struct Raw8<__nv_bfloat16> { uint4 r; };
__device__ __forceinline__ void ldraw(const float* __restrict__ p, Raw8<float>& v) {
    v.a = __ldg(reinterpret_cast<const float4*>(p));
    v.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
}
__device__ __forceinline__ void ldraw(const __nv_bfloat16* __restrict__ p, Raw8<__nv_bfloat16>& v) {
    v.r = __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void unpack8(const Raw8<float>& r, float v[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
__device__ __forceinline__ void unpack8(const Raw8<__nv_bfloat16>& r, float v[8]) {
    const uint32_t w[4] = {r.r.x, r.r.y, r.r.z, r.r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void store8(float* __restrict__ p, const float v[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* __restrict__ p, const float v[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float gelu_f(float z) {          // exact-erf GELU
    return 0.5f * z * (1.0f + erff(z * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_grad_f(float z) {     // d/dz [z * Phi(z)]
    const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * z * z);
    return cdf + z * pdf;
}
// Phi(z) and exp(-z^2/2) from ONE exponential: erf(x) = 1 - (a1 t + .. + a5 t^5) exp(-x^2), t = 1/(1 + p x), x = |z|/sqrt(2)
// (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7: far below bf16's 2^-9).  Used only where the result is rounded to bf16; the
// fp32 strict-parity path keeps erff.
__device__ __forceinline__ float gelu_cdf_fast(float z, float& e) {
    const float ax = fabsf(z) * 0.70710678118654752440f;
    const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
    e = __expf(-ax * ax);
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float half_tail = 0.5f * poly * t * e;            // 0.5 * (1 - erf(ax))
    return z >= 0.f ? 1.0f - half_tail : half_tail;
}
template <typename T> __device__ __forceinline__ float gelu_t(float z) { return gelu_f(z); }
template <> __device__ __forceinline__ float gelu_t<__nv_bfloat16>(float z) { float e; return z * gelu_cdf_fast(z, e); }
template <typename T> __device__ __forceinline__ float gelu_grad_t(float z) { return gelu_grad_f(z); }
template <> __device__ __forceinline__ float gelu_grad_t<__nv_bfloat16>(float z) {
    float e;
    const float cdf = gelu_cdf_fast(z, e);
    return fmaf(z * 0.39894228040143267794f, e, cdf);
}
__device__ __forceinline__ float sigmoid_f(float z) { return 1.0f / (1.0f + __expf(-z)); }

// (group, x, y, b) of a flat index over [b][y][x][group].  The element counts of this path fit 32 bits (B=16 x 128^2 x 60 groups =
// 15.7 M), and five 64-bit divisions per 16 bytes of output were most of the instructions of the gather / fuse kernels, so the
// 32-bit path is taken whenever the launch is small enough (`small` is launch-uniform).
struct PixIdx { int grp, x, y, b; int64_t pix; };
__device__ __forceinline__ PixIdx split_pix(int64_t idx, int groups, int W, int H, bool small) {
    PixIdx r;
    if (small) {
        const uint32_t i = (uint32_t)idx, pix = i / (uint32_t)groups, t = pix / (uint32_t)W;
        r.grp = (int)(i - pix * (uint32_t)groups);
        r.x = (int)(pix - t * (uint32_t)W);
        r.b = (int)(t / (uint32_t)H);
        r.y = (int)(t - (uint32_t)r.b * (uint32_t)H);
        r.pix = pix;
    } else {
        r.grp = (int)(idx % groups);
        r.pix = idx / groups;
        r.x = (int)(r.pix % W);
        r.y = (int)((r.pix / W) % H);
        r.b = (int)(r.pix / ((int64_t)W * H));
    }
    return r;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------------
// The step is ~3000 short kernels; inside the replayed CUDA graph a dependent kernel node starts ~2-3 us after its predecessor
// ends.  Kernels launched through launch_k() carry cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may become
// resident (and run their prologue: barrier init, TMEM allocation, weight staging) while the predecessor drains; pdl_wait() at
// the top of the kernel blocks until every prerequisite grid has completed and its writes are visible, so correctness never
// depends on the overlap.  pdl_trigger() lets the NEXT kernel's launch begin once all CTAs of this one are past it.
// Launched normally (the default, or a predecessor that is not a kernel) both instructions are no-ops.
// MEASURED (gpurun 2026-10-17, B=16 step, BatchNorm / LayerNorm / fuse / fused-conv kernels converted): 500.0 img/s with PDL vs
// 507.8 without -- the early-resident CTAs of the dependent kernel take SM slots away from the other streams of the step
// (average concurrency 1.9) and that costs more than the launch latency they hide.  OFF by default; RSS_PDL=1 enables it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("RSS_PDL");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on != 0;
}

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface through check_launch()
}

// Resident blocks per SM granted to the element-wise streaming kernels of the multi-stream part of the step (fuse sums, LayerNorm).
// A grid of 8 x 256-thread blocks per SM takes every thread slot, so kernels of the step's other streams cannot start beside it;
// see bn.cu bn_apply_bpsm() for the measurement that set the BatchNorm kernels to 4.  RSS_STREAM_BPSM overrides `dflt`.
inline int stream_bpsm(int dflt) {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RSS_STREAM_BPSM"); v = e ? atoi(e) : 0; if (v < 0 || v > 32) v = 0; }
    return v > 0 ? v : dflt;
}

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace rss
