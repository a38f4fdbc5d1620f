// Fusion neck gather, 1x1 classifier head and auxiliary head of RSSFormer (NHWC, HBM-bound kernels).
//
// Reference: RSSFormer-TIP2023/module/baseline/hrnet_aux.py
//   SimpleFusion8.forward :51-68  3x F.interpolate(bilinear, align_corners=True) + cat -> 480 channels
//   head :78-81                   Conv2d(480,7,1) + UpsamplingBilinear2d(x4) (== align_corners=True)
//   eval output :109-110          softmax(dim=1) of the up-sampled logits
//   headaux :86-87,99-101         AdaptiveAvgPool2d(1) + Linear(32,7)
// The concat is written once, directly in its final NHWC place (the reference writes three
// up-sampled tensors and then copies all four into the concat); the x4 up-sampling of the logits
// is never materialised in training (it is fused into the loss kernel, loss.cu).
#include "common.cuh"

namespace rss {

// PyTorch's align_corners=True source index (area_pixel_compute_source_index): src = o * (in-1)/(out-1)
__device__ __forceinline__ void bilinear_src(int o, float scale, int in, int& i0, int& i1, float& l1) {
    const float src = scale * (float)o;
    i0 = (int)src;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = src - (float)i0;
}
static inline float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

struct NeckGeom {
    int B, H, W;                 // output (level-0) size
    int C[4], h[4], w[4], coff[4];
    int Ctot;
    float sy[4], sx[4];
};

template <typename T>
__global__ void neck_gather_fwd_kernel(const T* __restrict__ f0, const T* __restrict__ f1, const T* __restrict__ f2,
                                       const T* __restrict__ f3, T* __restrict__ out, NeckGeom g) {
    const int groups = g.Ctot / 8;
    const int64_t total = (int64_t)g.B * g.H * g.W * groups;
    const bool small = total < 0x7fffffffLL;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const PixIdx q = split_pix(idx, groups, g.W, g.H, small);
        const int grp = q.grp, ox = q.x, oy = q.y, b = q.b;
        const int64_t pix = q.pix;
        const int ch = grp * 8;
        int l = 0;
        if (ch >= g.coff[3]) l = 3; else if (ch >= g.coff[2]) l = 2; else if (ch >= g.coff[1]) l = 1;
        const T* src = l == 0 ? f0 : (l == 1 ? f1 : (l == 2 ? f2 : f3));
        const int c = ch - g.coff[l], Cl = g.C[l], hl = g.h[l], wl = g.w[l];
        float v[8];
        if (l == 0) {
            load8(src + (((int64_t)b * hl + oy) * wl + ox) * Cl + c, v);
        } else {
            int y0, y1, x0, x1; float ly, lx;
            bilinear_src(oy, g.sy[l], hl, y0, y1, ly);
            bilinear_src(ox, g.sx[l], wl, x0, x1, lx);
            float a[8], bb[8], cc[8], d[8];
            const T* base = src + (int64_t)b * hl * wl * Cl + c;
            load8(base + ((int64_t)y0 * wl + x0) * Cl, a);
            load8(base + ((int64_t)y0 * wl + x1) * Cl, bb);
            load8(base + ((int64_t)y1 * wl + x0) * Cl, cc);
            load8(base + ((int64_t)y1 * wl + x1) * Cl, d);
            const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = hy * (hx * a[i] + lx * bb[i]) + ly * (hx * cc[i] + lx * d[i]);
        }
        store8(out + pix * g.Ctot + ch, v);
    }
}

// Gather-form transpose of the bilinear up-sampling, ALL FOUR levels in one launch (coarsest first: its items are the longest).
// One item = (input pixel, 8 channels); at the coarse levels an item is split over `parts` = 2 / 4 adjacent lanes that take every
// parts-th output row of the pixel's footprint and are combined with shuffles (level 3 has only 131 k items of ~17 x 17 loads each:
// one thread per item left 3.5 half-empty blocks per SM walking 441 candidates serially -- 136 us for a 134 MB read; the four
// launches ran back to back, 370 us of the one-stream tail of the step).
// align_corners=True makes the weight of output o on input cell i a tent, w = max(0, 1 - |min(s*o, in-1) - i|) -- the same numbers
// as bilinear_src() + the (i0 == i, i1 == i) tests of round 1 in 4 instructions instead of ~15 per candidate.
struct NeckBwdPlan { int blk_end[4]; int level[4]; int parts[4]; };

__device__ __forceinline__ float tent_weight(int o, float s, int in, int i) {
    return fmaxf(0.f, 1.f - fabsf(fminf(s * (float)o, (float)(in - 1)) - (float)i));
}

template <typename T>
__global__ void __launch_bounds__(256)
neck_gather_bwd_kernel(const T* __restrict__ dcat, T* __restrict__ d0, T* __restrict__ d1, T* __restrict__ d2, T* __restrict__ d3,
                       NeckGeom g, NeckBwdPlan pl) {
    int slot = 0;
    while (slot < 3 && (int)blockIdx.x >= pl.blk_end[slot]) ++slot;
    const int l = pl.level[slot], P = pl.parts[slot];
    const int Cl = g.C[l], hl = g.h[l], wl = g.w[l], groups = Cl / 8;
    T* dst = l == 0 ? d0 : (l == 1 ? d1 : (l == 2 ? d2 : d3));
    const int64_t total = (int64_t)g.B * hl * wl * groups;
    const int64_t item = ((int64_t)blockIdx.x - (slot ? pl.blk_end[slot - 1] : 0)) * blockDim.x + threadIdx.x;
    const int part = (int)(item & (P - 1));
    int64_t idx = item / P;
    const bool live = idx < total;                      // dead lanes keep running: they take part in the shuffles
    if (!live) idx = total - 1;
    const PixIdx q = split_pix(idx, groups, wl, hl, total < 0x7fffffffLL);
    const int grp = q.grp, ix = q.x, iy = q.y, b = q.b;
    const T* src = dcat + (int64_t)b * g.H * g.W * g.Ctot + g.coff[l] + grp * 8;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (l == 0) {
        if (live) load8(src + ((int64_t)iy * g.W + ix) * g.Ctot, acc);
    } else if (live) {
        const float sy = g.sy[l], sx = g.sx[l];
        int oy_lo = sy > 0.f ? (int)floorf((iy - 1) / sy) - 1 : 0, oy_hi = sy > 0.f ? (int)ceilf((iy + 1) / sy) + 1 : g.H - 1;
        int ox_lo = sx > 0.f ? (int)floorf((ix - 1) / sx) - 1 : 0, ox_hi = sx > 0.f ? (int)ceilf((ix + 1) / sx) + 1 : g.W - 1;
        if (oy_lo < 0) oy_lo = 0; if (ox_lo < 0) ox_lo = 0;
        if (oy_hi > g.H - 1) oy_hi = g.H - 1; if (ox_hi > g.W - 1) ox_hi = g.W - 1;
        while (ox_lo < ox_hi && tent_weight(ox_lo, sx, wl, ix) == 0.f) ++ox_lo;      // trim the safety margins: the inner loop is branch-free
        while (ox_hi > ox_lo && tent_weight(ox_hi, sx, wl, ix) == 0.f) --ox_hi;
        for (int oy = oy_lo + part; oy <= oy_hi; oy += P) {
            const float wy = tent_weight(oy, sy, hl, iy);
            if (wy == 0.f) continue;
            const T* rowp = src + (int64_t)oy * g.W * g.Ctot;
#pragma unroll 4
            for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                float v[8];
                load8(rowp + (int64_t)ox * g.Ctot, v);
                const float wgt = wy * tent_weight(ox, sx, wl, ix);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += wgt * v[i];
            }
        }
    }
    if (P > 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 1);
            if (P > 2) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 2);
        }
    }
    if (live && part == 0) store8(dst + q.pix * Cl + grp * 8, acc);
}

// ---------------------------------------------------------------------------------------------
// head: logits_lr[pix][o] = sum_c x[pix][c] * W[o][c] + b[o]   (o<7, padded to 8 floats per pixel)
// ---------------------------------------------------------------------------------------------
constexpr int kNC = 7, kNCP = 8;

// General-C version: one warp per pixel, weights in shared memory (14 LDS.128 per 16-byte activation chunk: shared-memory bound).
template <typename T>
__global__ void __launch_bounds__(256) head_fwd_smem_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ logits, int64_t pixels, int C) {
    extern __shared__ float wsm[];                    // [7][C]
    for (int i = threadIdx.x; i < kNC * C; i += blockDim.x) wsm[i] = w[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int chunks = C / 8;
    for (int64_t pix = (int64_t)blockIdx.x * wpb + warp; pix < pixels; pix += (int64_t)gridDim.x * wpb) {
        float acc[kNC];
#pragma unroll
        for (int o = 0; o < kNC; ++o) acc[o] = 0.f;
        for (int ck = lane; ck < chunks; ck += 32) {
            float v[8];
            load8(x + pix * C + ck * 8, v);
#pragma unroll
            for (int o = 0; o < kNC; ++o) {
                const float4 w0 = *reinterpret_cast<const float4*>(wsm + o * C + ck * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(wsm + o * C + ck * 8 + 4);
                acc[o] += v[0] * w0.x + v[1] * w0.y + v[2] * w0.z + v[3] * w0.w + v[4] * w1.x + v[5] * w1.y + v[6] * w1.z + v[7] * w1.w;
            }
        }
#pragma unroll
        for (int o = 0; o < kNC; ++o) acc[o] = warp_sum(acc[o]);
        if (lane < kNCP) {
            float r = 0.f;
#pragma unroll
            for (int o = 0; o < kNC; ++o) if (lane == o) r = acc[o] + bias[o];
            logits[pix * kNCP + lane] = r;
        }
    }
}

// C <= 512 (RSSFormer: 480): one warp per pixel, lane L owns the FIXED channel chunks L and L + 32 of every pixel, so its 2 x 7 x 8
// weights live in registers for the whole kernel (the shared-memory version above spends 14 LDS.128 per loaded chunk and ran at
// 0.15 of the HBM roofline).  kHeadU pixels per warp iteration keep 2 * kHeadU 16-byte loads in flight per lane.
constexpr int kHeadU = 4;
template <typename T>
__global__ void __launch_bounds__(128) head_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ logits, int64_t pixels, int C) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int chunks = C / 8;
    const bool has0 = lane < chunks, has1 = lane + 32 < chunks;
    float wr0[kNC][8], wr1[kNC][8];
#pragma unroll
    for (int o = 0; o < kNC; ++o)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            wr0[o][i] = has0 ? w[o * C + lane * 8 + i] : 0.f;
            wr1[o][i] = has1 ? w[o * C + (lane + 32) * 8 + i] : 0.f;
        }
    const float bl = lane < kNC ? bias[lane] : 0.f;
    const int64_t stride = (int64_t)gridDim.x * wpb;
    for (int64_t pix0 = (int64_t)blockIdx.x * wpb + warp; pix0 < pixels; pix0 += kHeadU * stride) {
        Raw8<T> r0[kHeadU], r1[kHeadU];
#pragma unroll
        for (int u = 0; u < kHeadU; ++u) {
            const int64_t pix = pix0 + u * stride;
            if (pix < pixels) {
                if (has0) ldraw(x + pix * C + lane * 8, r0[u]);
                if (has1) ldraw(x + pix * C + (lane + 32) * 8, r1[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < kHeadU; ++u) {
            const int64_t pix = pix0 + u * stride;
            if (pix >= pixels) break;                  // warp-uniform
            float acc[kNC];
#pragma unroll
            for (int o = 0; o < kNC; ++o) acc[o] = 0.f;
            if (has0) {
                float v[8];
                unpack8(r0[u], v);
#pragma unroll
                for (int o = 0; o < kNC; ++o)
                    acc[o] += v[0] * wr0[o][0] + v[1] * wr0[o][1] + v[2] * wr0[o][2] + v[3] * wr0[o][3] + v[4] * wr0[o][4] + v[5] * wr0[o][5] +
                              v[6] * wr0[o][6] + v[7] * wr0[o][7];
            }
            if (has1) {
                float v[8];
                unpack8(r1[u], v);
#pragma unroll
                for (int o = 0; o < kNC; ++o)
                    acc[o] += v[0] * wr1[o][0] + v[1] * wr1[o][1] + v[2] * wr1[o][2] + v[3] * wr1[o][3] + v[4] * wr1[o][4] + v[5] * wr1[o][5] +
                              v[6] * wr1[o][6] + v[7] * wr1[o][7];
            }
#pragma unroll
            for (int o = 0; o < kNC; ++o) acc[o] = warp_sum(acc[o]);
            if (lane < kNCP) {
                float r = 0.f;
#pragma unroll
                for (int o = 0; o < kNC; ++o) if (lane == o) r = acc[o] + bl;
                logits[pix * kNCP + lane] = r;
            }
        }
    }
}

// (A split into a dx pass and a dW pass with ~56 state registers each was measured slower on B200 -- 105 + 160 us vs 218 us for
//  this fused kernel: both passes stayed latency-bound on their one-pixel-per-iteration loops -- so the fused kernel stays.)
// dx[pix][c] = sum_o dl[pix][o] W[o][c];  dW[o][c] += sum_pix dl[pix][o] x[pix][c];  db[o] += sum_pix dl[pix][o]
template <typename T>
__global__ void head_bwd_kernel(const T* __restrict__ x, const float* __restrict__ dl, const float* __restrict__ w,
                                T* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db,
                                int64_t pixels, int C, int cg, int rpb) {
    extern __shared__ float red[];                    // [rpb][7][C]
    const int sub = threadIdx.x % cg, r = threadIdx.x / cg;
    float wr[kNC][8], aw[kNC][8], ab[kNC];
#pragma unroll
    for (int o = 0; o < kNC; ++o) {
        ab[o] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { wr[o][i] = w[o * C + sub * 8 + i]; aw[o][i] = 0.f; }
    }
    for (int64_t pix = (int64_t)blockIdx.x * rpb + r; pix < pixels; pix += (int64_t)gridDim.x * rpb) {
        const float4 d0 = __ldg(reinterpret_cast<const float4*>(dl + pix * kNCP));
        const float4 d1 = __ldg(reinterpret_cast<const float4*>(dl + pix * kNCP) + 1);
        const float d[kNC] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z};
        float v[8], o8[8];
        load8(x + pix * C + sub * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) o8[i] = 0.f;
#pragma unroll
        for (int o = 0; o < kNC; ++o) {
            if (sub == 0) ab[o] += d[o];
#pragma unroll
            for (int i = 0; i < 8; ++i) { o8[i] += d[o] * wr[o][i]; aw[o][i] += d[o] * v[i]; }
        }
        store8(dx + pix * C + sub * 8, o8);
    }
#pragma unroll
    for (int o = 0; o < kNC; ++o)
#pragma unroll
        for (int i = 0; i < 8; ++i) red[((size_t)r * kNC + o) * C + sub * 8 + i] = aw[o][i];
    __syncthreads();
    for (int e = threadIdx.x; e < kNC * C; e += blockDim.x) {
        float s = 0.f;
        for (int rr = 0; rr < rpb; ++rr) s += red[(size_t)rr * kNC * C + e];
        atomicAdd(dw + e, s);
    }
    __syncthreads();
    if (sub == 0) {
#pragma unroll
        for (int o = 0; o < kNC; ++o) red[r * kNC + o] = ab[o];
    }
    __syncthreads();
    if (threadIdx.x < kNC) {
        float s = 0.f;
        for (int rr = 0; rr < rpb; ++rr) s += red[rr * kNC + threadIdx.x];
        atomicAdd(db + threadIdx.x, s);
    }
}

// eval: probs[b][c][oy][ox] = softmax_c(bilinear_x4(logits_lr))   (NCHW fp32, as the reference returns), argmax optional
__global__ void head_probs_kernel(const float* __restrict__ logits, float* __restrict__ probs, uint8_t* __restrict__ argmax,
                                  int B, int h, int w, int H, int W, float sy, float sx) {
    const int64_t total = (int64_t)B * H * W;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int ox = (int)(idx % W), oy = (int)((idx / W) % H), b = (int)(idx / ((int64_t)W * H));
        int y0, y1, x0, x1; float ly, lx;
        bilinear_src(oy, sy, h, y0, y1, ly);
        bilinear_src(ox, sx, w, x0, x1, lx);
        const float* base = logits + (int64_t)b * h * w * kNCP;
        const float* p00 = base + ((int64_t)y0 * w + x0) * kNCP; const float* p01 = base + ((int64_t)y0 * w + x1) * kNCP;
        const float* p10 = base + ((int64_t)y1 * w + x0) * kNCP; const float* p11 = base + ((int64_t)y1 * w + x1) * kNCP;
        const float hy = 1.f - ly, hx = 1.f - lx;
        float z[kNC], mx = -INFINITY;
        int am = 0;
#pragma unroll
        for (int c = 0; c < kNC; ++c) {
            z[c] = hy * (hx * p00[c] + lx * p01[c]) + ly * (hx * p10[c] + lx * p11[c]);
            if (z[c] > mx) { mx = z[c]; am = c; }
        }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < kNC; ++c) { z[c] = expf(z[c] - mx); s += z[c]; }
        const float inv = 1.f / s;
        const int64_t plane = (int64_t)H * W;
#pragma unroll
        for (int c = 0; c < kNC; ++c) probs[((int64_t)b * kNC + c) * plane + (int64_t)oy * W + ox] = z[c] * inv;
        if (argmax) argmax[idx] = (uint8_t)am;
    }
}

// headaux: per-image channel sums of f0, then Linear(C -> 7)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, float* __restrict__ sums /*[B][C]*/, int HW, int C, int cg, int rpb) {
    extern __shared__ float red[];                    // [rpb][C]
    const int b = blockIdx.y, sub = threadIdx.x % cg, r = threadIdx.x / cg;
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.f;
    for (int row = blockIdx.x * rpb + r; row < HW; row += gridDim.x * rpb) {
        float v[8];
        load8(x + ((int64_t)b * HW + row) * C + sub * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[r * C + sub * 8 + i] = a[i];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int rr = 0; rr < rpb; ++rr) s += red[rr * C + c];
        atomicAdd(sums + b * C + c, s);
    }
}

__global__ void headaux_linear_kernel(const float* __restrict__ sums, const float* __restrict__ w, const float* __restrict__ bias,
                                      float* __restrict__ scores, int B, int C, float inv_hw) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * kNC) return;
    const int b = idx / kNC, o = idx % kNC;
    float s = bias[o];
    for (int c = 0; c < C; ++c) s += w[o * C + c] * (sums[b * C + c] * inv_hw);
    scores[idx] = s;
}

static NeckGeom neck_geom(int B, int H, int W, const int* C, const int* h, const int* w) {
    NeckGeom g; g.B = B; g.H = H; g.W = W; g.Ctot = 0;
    for (int l = 0; l < 4; ++l) {
        g.C[l] = C[l]; g.h[l] = h[l]; g.w[l] = w[l]; g.coff[l] = g.Ctot; g.Ctot += C[l];
        g.sy[l] = ac_scale(h[l], H); g.sx[l] = ac_scale(w[l], W);
    }
    return g;
}

}  // namespace rss

using namespace rss;

extern "C" int rss_neck_gather_fwd(const void* f0, const void* f1, const void* f2, const void* f3, void* out,
                                   int B, const int* C, const int* h, const int* w, int dtype, cudaStream_t st) {
    if (B <= 0 || !C || !h || !w) return RSS_ERR_SHAPE;
    for (int l = 0; l < 4; ++l) if (C[l] <= 0 || C[l] % 8 || h[l] <= 0 || w[l] <= 0) return RSS_ERR_SHAPE;
    const NeckGeom g = neck_geom(B, h[0], w[0], C, h, w);
    const int64_t total = (int64_t)B * g.H * g.W * (g.Ctot / 8);
    int grid = (int)((total + 255) / 256);
    if (grid > num_sms() * 16) grid = num_sms() * 16;
    RSS_DISPATCH_DTYPE(dtype, neck_gather_fwd_kernel<T><<<grid, 256, 0, st>>>((const T*)f0, (const T*)f1, (const T*)f2, (const T*)f3, (T*)out, g));
    return check_launch();
}

extern "C" int rss_neck_gather_bwd(const void* dcat, void* d0, void* d1, void* d2, void* d3,
                                   int B, const int* C, const int* h, const int* w, int dtype, cudaStream_t st) {
    if (B <= 0 || !C || !h || !w) return RSS_ERR_SHAPE;
    for (int l = 0; l < 4; ++l) if (C[l] <= 0 || C[l] % 8 || h[l] <= 0 || w[l] <= 0) return RSS_ERR_SHAPE;
    const NeckGeom g = neck_geom(B, h[0], w[0], C, h, w);
    NeckBwdPlan pl;
    int64_t blocks = 0;
    for (int s = 0; s < 4; ++s) {                        // coarsest level first
        const int l = 3 - s;
        const int ratio = h[l] > 0 ? h[0] / h[l] : 1;
        pl.level[s] = l;
        pl.parts[s] = l == 0 ? 1 : (ratio >= 8 ? 4 : (ratio >= 4 ? 2 : 1));
        const int64_t items = (int64_t)B * h[l] * w[l] * (C[l] / 8) * pl.parts[s];
        blocks += (items + 255) / 256;
        if (blocks > 0x7fffffff) return RSS_ERR_SHAPE;
        pl.blk_end[s] = (int)blocks;
    }
    RSS_DISPATCH_DTYPE(dtype, neck_gather_bwd_kernel<T><<<(int)blocks, 256, 0, st>>>((const T*)dcat, (T*)d0, (T*)d1, (T*)d2, (T*)d3, g, pl));
    return check_launch();
}

extern "C" int rss_head_fwd(const void* x, const float* w, const float* bias, float* logits_lr, int64_t pixels, int C,
                            int dtype, cudaStream_t st) {
    if (pixels <= 0 || C <= 0 || C % 8 || (size_t)kNC * C * sizeof(float) > 48 * 1024) return RSS_ERR_SHAPE;
    if (C <= 512) {                                   // weights in registers (2 chunks per lane)
        int grid = (int)((pixels + 3) / 4);
        if (grid > num_sms() * 2) grid = num_sms() * 2;      // 181 registers x 128 threads: two resident blocks per SM
        RSS_DISPATCH_DTYPE(dtype, head_fwd_kernel<T><<<grid, 128, 0, st>>>((const T*)x, w, bias, logits_lr, pixels, C));
        return check_launch();
    }
    int grid = (int)((pixels + 7) / 8);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    RSS_DISPATCH_DTYPE(dtype, head_fwd_smem_kernel<T><<<grid, 256, kNC * C * sizeof(float), st>>>((const T*)x, w, bias, logits_lr, pixels, C));
    return check_launch();
}

extern "C" int rss_head_bwd(const void* x, const float* dlogits_lr, const float* w, void* dx, float* dw_acc, float* db_acc,
                            int64_t pixels, int C, int dtype, cudaStream_t st) {
    if (pixels <= 0 || C <= 0 || C % 8) return RSS_ERR_SHAPE;
    const int cg = C / 8;
    // rows per block: RSS_HEAD_BWD_RPB (default 4).  The kernel holds 168 registers, so a 240-thread block (4 rows at C = 480) is alone
    // on its SM; measured on the B=16 step: 120-thread blocks, three per SM, 247 us vs 219 us (1 row: slower still) -- the per-block
    // shared-memory reduction + 3360 atomics of the weight gradient cost more than the extra warps hide.  A register double buffer
    // of the next pixel's operands was measured at 252 us, two pixels per iteration with all loads issued first at 315 us.  Kept at 4.
    static int rpb_max = 0;
    if (rpb_max == 0) { const char* e = getenv("RSS_HEAD_BWD_RPB"); rpb_max = e ? atoi(e) : 4; if (rpb_max < 1 || rpb_max > 4) rpb_max = 4; }
    int rpb = 256 / cg; if (rpb < 1) rpb = 1; if (rpb > rpb_max) rpb = rpb_max;
    while (rpb > 1 && (size_t)rpb * kNC * C * sizeof(float) > 48 * 1024) --rpb;
    const size_t smem = (size_t)rpb * kNC * C * sizeof(float);
    if (smem > 48 * 1024) return RSS_ERR_SHAPE;
    int grid = (int)((pixels + rpb - 1) / rpb);
    const int per_sm = 16 / rpb > 4 ? 16 / rpb : 4;
    if (grid > num_sms() * per_sm) grid = num_sms() * per_sm;
    RSS_DISPATCH_DTYPE(dtype, head_bwd_kernel<T><<<grid, cg * rpb, smem, st>>>((const T*)x, dlogits_lr, w, (T*)dx, dw_acc, db_acc, pixels, C, cg, rpb));
    return check_launch();
}

extern "C" int rss_head_probs(const float* logits_lr, float* probs, uint8_t* argmax, int B, int h, int w, int scale,
                              cudaStream_t st) {
    if (B <= 0 || h <= 0 || w <= 0 || scale <= 0) return RSS_ERR_SHAPE;
    const int H = h * scale, W = w * scale;
    const int64_t total = (int64_t)B * H * W;
    int grid = (int)((total + 255) / 256);
    if (grid > num_sms() * 16) grid = num_sms() * 16;
    head_probs_kernel<<<grid, 256, 0, st>>>(logits_lr, probs, argmax, B, h, w, H, W, ac_scale(h, H), ac_scale(w, W));
    return check_launch();
}

extern "C" int rss_headaux_fwd(const void* f0, const float* w, const float* bias, float* colsum_ws, float* scores,
                               int B, int HW, int C, int dtype, cudaStream_t st) {
    if (B <= 0 || HW <= 0 || C <= 0 || C % 8 || C > 1024) return RSS_ERR_SHAPE;
    cudaError_t e = cudaMemsetAsync(colsum_ws, 0, (size_t)B * C * sizeof(float), st);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return RSS_ERR_CUDA; }
    const int cg = C / 8;
    int rpb = 256 / cg; if (rpb < 1) rpb = 1;
    int gx = (HW + rpb * 16 - 1) / (rpb * 16); if (gx < 1) gx = 1; if (gx > 64) gx = 64;
    dim3 grid(gx, B);
    RSS_DISPATCH_DTYPE(dtype, colsum_kernel<T><<<grid, cg * rpb, (size_t)rpb * C * sizeof(float), st>>>((const T*)f0, colsum_ws, HW, C, cg, rpb));
    headaux_linear_kernel<<<(B * kNC + 127) / 128, 128, 0, st>>>(colsum_ws, w, bias, scores, B, C, 1.0f / (float)HW);
    return check_launch();
}
