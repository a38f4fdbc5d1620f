// LayerNorm over the channel axis of NHWC tokens (reference: modules/MTFM.py:64,78-79,107,109; eps 1e-6).
// HBM-bound: one 16-byte (bf16) / 32-byte (fp32) vector per thread, C/8 lanes cooperate on a token
// through warp shuffles; fp32 statistics.
#include "common.cuh"

namespace rss {

template <int TPT>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = TPT / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T, int TPT>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float eps, int64_t rows) {
    pdl_wait();
    pdl_trigger();
    constexpr int C = TPT * 8;
    const int sub = threadIdx.x % TPT;
    float g[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = gamma[sub * 8 + i]; b[i] = beta[sub * 8 + i]; }
    const int64_t rows_per_block = blockDim.x / TPT;
    // every lane of a warp iterates the same number of times (shuffles need full participation)
    const int64_t iters = (rows + rows_per_block * gridDim.x - 1) / (rows_per_block * gridDim.x);
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t row = (it * gridDim.x + blockIdx.x) * rows_per_block + threadIdx.x / TPT;
        const bool live = row < rows;
        float v[8];
        if (live) load8(x + row * C + sub * 8, v);
        else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
        const float mu = group_sum<TPT>(s) * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[i] - mu; q += d * d; }
        const float rstd = rsqrtf(group_sum<TPT>(q) * (1.0f / C) + eps);
        if (live) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = (v[i] - mu) * rstd * g[i] + b[i];
            if (y) store8(y + row * C + sub * 8, o);      // y == NULL: statistics only (consumers re-normalise on the fly)
            if (sub == 0 && mean_out) { mean_out[row] = mu; rstd_out[row] = rstd; }
        }
    }
}

// dx = rstd * (dy*g - mean_c(dy*g) - xhat * mean_c(dy*g*xhat)) [+ dx_add];  dgamma += sum_rows dy*xhat; dbeta += sum_rows dy
template <typename T, typename TDY, int TPT>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const TDY* __restrict__ dy, const T* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const T* __restrict__ dx_add,
                                                     T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     int64_t rows) {
    pdl_wait();
    pdl_trigger();
    constexpr int C = TPT * 8;
    __shared__ float red[2][256 / TPT][C + 1];
    const int sub = threadIdx.x % TPT;
    const int grp = threadIdx.x / TPT;
    float g[8], ag[8], ab[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = gamma[sub * 8 + i]; ag[i] = 0.f; ab[i] = 0.f; }
    const int64_t rows_per_block = blockDim.x / TPT;
    const int64_t iters = (rows + rows_per_block * gridDim.x - 1) / (rows_per_block * gridDim.x);
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t row = (it * gridDim.x + blockIdx.x) * rows_per_block + grp;
        const bool live = row < rows;
        float d[8], v[8];
        float mu = 0.f, rs = 0.f;
        if (live) {
            load8(dy + row * C + sub * 8, d);
            load8(x + row * C + sub * 8, v);
            mu = mean[row]; rs = rstd[row];
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) { d[i] = 0.f; v[i] = 0.f; }
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = (v[i] - mu) * rs;                  // xhat
            ag[i] += d[i] * v[i];
            ab[i] += d[i];
            d[i] *= g[i];                             // dy*gamma
            s1 += d[i];
            s2 += d[i] * v[i];
        }
        s1 = group_sum<TPT>(s1) * (1.0f / C);
        s2 = group_sum<TPT>(s2) * (1.0f / C);
        if (live) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = rs * (d[i] - s1 - v[i] * s2);
            if (dx_add) {
                float a[8];
                load8(dx_add + row * C + sub * 8, a);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] += a[i];
            }
            store8(dx + row * C + sub * 8, o);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[0][grp][sub * 8 + i] = ag[i]; red[1][grp][sub * 8 + i] = ab[i]; }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
        const int which = c / C, ch = c % C;
        float s = 0.f;
        for (int r = 0; r < (int)rows_per_block; ++r) s += red[which][r][ch];
        atomicAdd((which == 0 ? dgamma : dbeta) + ch, s);
    }
}

// LayerNorm-1 backward of BOTH token streams of the attention region (x: queries + residual, y: keys/values) in one launch, with the
// element-wise tail of the saliency-gate backward applied to the incoming gradient on load:
//   d normed[f] = d gated[f] * gate[f mod HW] + d avgpool[f mod HW] / C + (argmax[f mod HW] == f div HW ? d maxpool[f mod HW] : 0)
// (f = flat index n*C + c of the token tensor of one image: the reference pools over the flat (C, H, W) VIEW of token memory,
// multihead_isa_pool_attention.py:150-151).  Replaces gate_bwd_apply_kernel (an in-place fp32 read-modify-write of both 33.5 MB
// gradient tensors) + two ln_bwd_kernel launches; same expressions in the same order, so the results are bit-identical.
// Needs HW % 8 == 0 (the 8 flat positions of a lane stay inside one row of the view); C = 32.
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_gated_kernel(const float* __restrict__ dxg, const float* __restrict__ dyg, const T* __restrict__ x,
                                                           const T* __restrict__ y, const float* __restrict__ stats /*[4][rows]*/,
                                                           const float* __restrict__ gamma, const T* __restrict__ dx_add,
                                                           T* __restrict__ dx, T* __restrict__ dy, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int64_t rows, int HW,
                                                           const float* __restrict__ gmap, const float* __restrict__ dpooled,
                                                           const uint8_t* __restrict__ amax) {
    constexpr int TPT = 4, C = 32;
    __shared__ float red[2][256 / TPT][C + 1];
    const int z = blockIdx.y;
    const float* dyv = z == 0 ? dxg : dyg;
    const T* xv = z == 0 ? x : y;
    T* dxo = z == 0 ? dx : dy;
    const T* add = z == 0 ? dx_add : nullptr;
    const float* mean = stats + (size_t)(2 * z) * rows;
    const float* rstd = stats + (size_t)(2 * z + 1) * rows;
    const int sub = threadIdx.x % TPT;
    const int grp = threadIdx.x / TPT;
    float g[8], ag[8], ab[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = gamma[sub * 8 + i]; ag[i] = 0.f; ab[i] = 0.f; }
    const int64_t rows_per_block = blockDim.x / TPT;
    const int64_t iters = (rows + rows_per_block * gridDim.x - 1) / (rows_per_block * gridDim.x);
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t row = (it * gridDim.x + blockIdx.x) * rows_per_block + grp;
        const bool live = row < rows;
        float d[8], v[8];
        float mu = 0.f, rs = 0.f;
        if (live) {
            load8(dyv + row * C + sub * 8, d);
            load8(xv + row * C + sub * 8, v);
            mu = mean[row]; rs = rstd[row];
            const int b = (int)(row / HW), n = (int)(row - (int64_t)b * HW);
            const uint32_t f0 = (uint32_t)n * C + sub * 8;
            const int kk = (int)(f0 / (uint32_t)HW), j = (int)(f0 - (uint32_t)kk * (uint32_t)HW);
            float g8[8], a8[8], m8[8];
            load8(gmap + ((size_t)b * 2 + z) * HW + j, g8);
            load8(dpooled + ((size_t)b * 4 + z * 2 + 0) * HW + j, a8);
            load8(dpooled + ((size_t)b * 4 + z * 2 + 1) * HW + j, m8);
            const uint2 am8 = __ldg(reinterpret_cast<const uint2*>(amax + ((size_t)b * 2 + z) * HW + j));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int a = (int)(((i < 4 ? am8.x : am8.y) >> (8 * (i & 3))) & 0xffu);
                d[i] = d[i] * g8[i] + a8[i] * (1.0f / C) + (a == kk ? m8[i] : 0.f);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) { d[i] = 0.f; v[i] = 0.f; }
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = (v[i] - mu) * rs;                  // xhat
            ag[i] += d[i] * v[i];
            ab[i] += d[i];
            d[i] *= g[i];                             // dy*gamma
            s1 += d[i];
            s2 += d[i] * v[i];
        }
        s1 = group_sum<TPT>(s1) * (1.0f / C);
        s2 = group_sum<TPT>(s2) * (1.0f / C);
        if (live) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = rs * (d[i] - s1 - v[i] * s2);
            if (add) {
                float a[8];
                load8(add + row * C + sub * 8, a);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] += a[i];
            }
            store8(dxo + row * C + sub * 8, o);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[0][grp][sub * 8 + i] = ag[i]; red[1][grp][sub * 8 + i] = ab[i]; }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
        const int which = c / C, ch = c % C;
        float s = 0.f;
        for (int r = 0; r < (int)rows_per_block; ++r) s += red[which][r][ch];
        atomicAdd((which == 0 ? dgamma : dbeta) + ch, s);
    }
}

template <typename T>
static int ln_fwd_launch(const void* x, void* y, float* mean, float* rstd, const float* gamma, const float* beta, float eps,
                         int64_t rows, int C, cudaStream_t st) {
    const int tpt = C / 8;
    const int64_t rpb = 256 / tpt;
    int grid = (int)((rows + rpb - 1) / rpb);
    const int cap = num_sms() * stream_bpsm(16);
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
#define LN_CASE(TPT) case TPT: launch_k(ln_fwd_kernel<T, TPT>, grid, 256, 0, st, (const T*)x, (T*)y, mean, rstd, gamma, beta, eps, rows); break;
    switch (tpt) { LN_CASE(1) LN_CASE(2) LN_CASE(4) LN_CASE(8) LN_CASE(16) LN_CASE(32) default: return RSS_ERR_SHAPE; }
#undef LN_CASE
    return check_launch();
}

template <typename T, typename TDY>
static int ln_bwd_launch(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                         const void* dx_add, void* dx, float* dgamma, float* dbeta, int64_t rows, int C, cudaStream_t st) {
    const int tpt = C / 8;
    const int64_t rpb = 256 / tpt;
    int grid = (int)((rows + rpb - 1) / rpb);
    const int cap = num_sms() * stream_bpsm(8);
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
#define LN_CASE(TPT) case TPT: launch_k(ln_bwd_kernel<T, TDY, TPT>, grid, 256, 0, st, (const TDY*)dy, (const T*)x, mean, rstd, gamma, (const T*)dx_add, (T*)dx, dgamma, dbeta, rows); break;
    switch (tpt) { LN_CASE(1) LN_CASE(2) LN_CASE(4) LN_CASE(8) LN_CASE(16) LN_CASE(32) default: return RSS_ERR_SHAPE; }
#undef LN_CASE
    return check_launch();
}

}  // namespace rss

extern "C" int rss_layernorm_fwd(const void* x, void* y, float* mean, float* rstd, const float* gamma, const float* beta,
                                 float eps, int64_t rows, int C, int dtype, cudaStream_t stream) {
    if (rows < 0 || C <= 0 || C % 8 != 0 || C > 256 || ((C / 8) & (C / 8 - 1))) return RSS_ERR_SHAPE;
    if (rows == 0) return RSS_OK;
    RSS_DISPATCH_DTYPE(dtype, return rss::ln_fwd_launch<T>(x, y, mean, rstd, gamma, beta, eps, rows, C, stream));
}

extern "C" int rss_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                                 const void* dx_add, void* dx, float* dgamma_acc, float* dbeta_acc,
                                 int64_t rows, int C, int dtype, cudaStream_t stream) {
    if (rows < 0 || C <= 0 || C % 8 != 0 || C > 256 || ((C / 8) & (C / 8 - 1))) return RSS_ERR_SHAPE;
    if (rows == 0) return RSS_OK;
    RSS_DISPATCH_DTYPE(dtype, return rss::ln_bwd_launch<T, T>(dy, x, mean, rstd, gamma, dx_add, dx, dgamma_acc, dbeta_acc, rows, C, stream));
}

// internal variant used by the attention region: the incoming gradient is fp32 whatever the activation dtype
namespace rss {
int layernorm_bwd_gated(const float* dxg, const float* dyg, const void* x, const void* y, const float* stats, const float* gamma,
                        const void* dx_add, void* dx, void* dy, float* dgamma_acc, float* dbeta_acc, int64_t rows, int HW,
                        const float* gmap, const float* dpooled, const uint8_t* amax, int dtype, cudaStream_t stream) {
    if (rows <= 0) return RSS_OK;
    if (HW <= 0 || (HW & 7) || rows % HW) return RSS_ERR_SHAPE;
    int grid = (int)((rows + 63) / 64);
    const int cap = num_sms() * 4;                       // x 2 token streams in grid.y
    if (grid > cap) grid = cap;
    RSS_DISPATCH_DTYPE(dtype, (ln_bwd_gated_kernel<T><<<dim3(grid, 2), 256, 0, stream>>>(dxg, dyg, (const T*)x, (const T*)y, stats, gamma,
                                                                                         (const T*)dx_add, (T*)dx, (T*)dy, dgamma_acc, dbeta_acc,
                                                                                         rows, HW, gmap, dpooled, amax)));
    return check_launch();
}
int layernorm_bwd_f32dy(const float* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                        const void* dx_add, void* dx, float* dgamma_acc, float* dbeta_acc, int64_t rows, int C, int dtype,
                        cudaStream_t stream) {
    if (rows <= 0) return RSS_OK;
    RSS_DISPATCH_DTYPE(dtype, return ln_bwd_launch<T, float>(dy, x, mean, rstd, gamma, dx_add, dx, dgamma_acc, dbeta_acc, rows, C, stream));
}
}
