// Multi-resolution fuse sum of HighResolutionModule.forward (_hrnet_rssformer.py:418-435):
//     out = [relu]( sum_j nearest_upsample_{2^k_j}(term_j) )
// and the residual + ReLU that closes a transformer block (MTFM.py:109 + _hrnet_rssformer.py:435) — one pass instead of
// (#terms) up-sample + add + relu kernels.  NHWC, every term has C channels and spatial (H >> k_j, W >> k_j).
#include "common.cuh"

namespace rss {

struct FuseTerms { const void* p[4]; int k[4]; int n; };

template <typename T>
__global__ void fuse_sum_fwd_kernel(FuseTerms t, T* __restrict__ out, int B, int H, int W, int C, int relu) {
    pdl_wait();
    pdl_trigger();
    const int groups = C / 8;
    const int64_t total = (int64_t)B * H * W * groups;
    const bool small = total < 0x7fffffffLL;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const PixIdx q = split_pix(idx, groups, W, H, small);
        const int grp = q.grp, x = q.x, y = q.y, b = q.b;
        const int64_t pix = q.pix;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < t.n) {
                const int k = t.k[j], h = H >> k, w = W >> k;
                float v[8];
                load8(reinterpret_cast<const T*>(t.p[j]) + (((int64_t)b * h + (y >> k)) * w + (x >> k)) * C + grp * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += v[i];
            }
        }
        if (relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaxf(acc[i], 0.f);
        }
        store8(out + pix * C + grp * 8, acc);
    }
}

// d term (at resolution H>>k) = sum over its 2^k x 2^k footprint of dout * [out > 0 if relu]
template <typename T>
__global__ void fuse_sum_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ out, T* __restrict__ dterm,
                                    int B, int H, int W, int C, int k, int relu) {
    pdl_wait();
    pdl_trigger();
    const int groups = C / 8, h = H >> k, w = W >> k, s = 1 << k;
    const int64_t total = (int64_t)B * h * w * groups;
    const bool small = total < 0x7fffffffLL;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const PixIdx q = split_pix(idx, groups, w, h, small);
        const int grp = q.grp, x = q.x, y = q.y, b = q.b;
        const int64_t pix = q.pix;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int dy = 0; dy < s; ++dy)
            for (int dx = 0; dx < s; ++dx) {
                const int64_t off = (((int64_t)b * H + (y * s + dy)) * W + (x * s + dx)) * C + grp * 8;
                float d[8];
                load8(dout + off, d);
                if (relu) {
                    float o[8];
                    load8(out + off, o);
#pragma unroll
                    for (int i = 0; i < 8; ++i) d[i] = o[i] > 0.f ? d[i] : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += d[i];
            }
        store8(dterm + pix * C + grp * 8, acc);
    }
}

}  // namespace rss

using namespace rss;

extern "C" int rss_fuse_sum_fwd(const void* const* terms, const int* log2_up, int n_terms, void* out, int B, int H, int W, int C,
                                int relu, int dtype, cudaStream_t st) {
    if (n_terms < 1 || n_terms > 4 || B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8) return RSS_ERR_SHAPE;
    FuseTerms t; t.n = n_terms;
    for (int j = 0; j < 4; ++j) { t.p[j] = j < n_terms ? terms[j] : nullptr; t.k[j] = j < n_terms ? log2_up[j] : 0; }
    for (int j = 0; j < n_terms; ++j) if (t.k[j] < 0 || t.k[j] > 5 || (H & ((1 << t.k[j]) - 1)) || (W & ((1 << t.k[j]) - 1))) return RSS_ERR_SHAPE;
    const int64_t total = (int64_t)B * H * W * (C / 8);
    int grid = (int)((total + 255) / 256);
    if (grid > num_sms() * stream_bpsm(16)) grid = num_sms() * stream_bpsm(16);
    RSS_DISPATCH_DTYPE(dtype, launch_k(fuse_sum_fwd_kernel<T>, grid, 256, 0, st, t, (T*)out, B, H, W, C, relu));
    return check_launch();
}

extern "C" int rss_fuse_sum_bwd(const void* dout, const void* out, void* dterm, int log2_up, int B, int H, int W, int C, int relu,
                                int dtype, cudaStream_t st) {
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 || log2_up < 0 || log2_up > 5) return RSS_ERR_SHAPE;
    if (relu && !out) return RSS_ERR_SHAPE;
    const int64_t total = (int64_t)B * (H >> log2_up) * (W >> log2_up) * (C / 8);
    int grid = (int)((total + 255) / 256);
    if (grid > num_sms() * stream_bpsm(16)) grid = num_sms() * stream_bpsm(16);
    if (grid < 1) grid = 1;
    RSS_DISPATCH_DTYPE(dtype, launch_k(fuse_sum_bwd_kernel<T>, grid, 256, 0, st, (const T*)dout, (const T*)out, (T*)dterm, B, H, W, C, log2_up, relu));
    return check_launch();
}
