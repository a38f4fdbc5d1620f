// Fused gate + 7x7 window cross-attention region of the RSSFormer transformer block.
//
// Reference semantics (paths under RSSFormer-TIP2023/module/baseline/base_hrnet/modules/):
//   LayerNorm1 on both inputs ................. MTFM.py:107 (same norm1 for x and y)
//   saliency gate on the re-interpreted view ... multihead_isa_pool_attention.py:148-167
//   centre zero-pad + 7x7 window gather ........ multihead_isa_attention.py:373-426 (no mask: pad tokens are live keys)
//   Mhca: q/k/v proj, softmax(QK^T)V, sigmoid(mean+max(Q^T K)) gate, out proj .. DAL.py:873-1020
//   residual add ............................... MTFM.py:107
//
// Data layout: tokens (B, H*W, 32) == NHWC.  One CTA owns one window at a time (persistent loop);
// the window's 49 tokens, q/k/v and the four 32x32 weights live in shared memory (row stride 36
// floats keeps every 128-bit access conflict-free); each attention row is owned by ONE thread so
// the softmax needs no cross-lane traffic and S/P never leave registers (the reference
// materialises (722*B,49,49) fp32 scores in HBM).  The region is HBM/issue-bound, not tensor-bound
// (49x16 tiles, K=16): see DESIGN.md.
#include "common.cuh"

namespace rss {

constexpr int kC = 32, kHD = 16, kL = 49, kWS = 7, kLD = 36, kPS = 49;
constexpr int kTok = kL * kLD;                 // floats per token buffer
constexpr int kThreads = 128;

struct WinGeom { int B, H, W, HW, ph, pw, qh, qw, nWin; };

int layernorm_bwd_f32dy(const float* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                        const void* dx_add, void* dx, float* dgamma_acc, float* dbeta_acc, int64_t rows, int C, int dtype,
                        cudaStream_t stream);
int layernorm_bwd_gated(const float* dxg, const float* dyg, const void* x, const void* y, const float* stats, const float* gamma,
                        const void* dx_add, void* dx, void* dy, float* dgamma_acc, float* dbeta_acc, int64_t rows, int HW,
                        const float* gmap, const float* dpooled, const uint8_t* amax, int dtype, cudaStream_t stream);

static inline WinGeom make_geom(int B, int H, int W) {
    WinGeom g; g.B = B; g.H = H; g.W = W; g.HW = H * W;
    const int Hp = (H + kWS - 1) / kWS * kWS, Wp = (W + kWS - 1) / kWS * kWS;
    g.ph = (Hp - H) / 2; g.pw = (Wp - W) / 2; g.qh = Hp / kWS; g.qw = Wp / kWS; g.nWin = B * g.qh * g.qw;
    return g;
}

// ------------------------------------------------------------------------------------------
// gate: pooled maps over the flat view, 7x7 conv + sigmoid, 1x1 conv + softmax
// ------------------------------------------------------------------------------------------
// LayerNorm applied on the fly: the normalised tokens are never written to HBM, and every consumer sees the
// fp32 value (x - mean)*rstd*gamma + beta whatever the storage dtype of x (ln == NULL: x is used as is).
struct LnRef { const float* mean; const float* rstd; const float* gamma; const float* beta; };

template <typename T>
__device__ __forceinline__ float normed_at(const T* __restrict__ base, const LnRef& ln, size_t img_row0, size_t f) {
    float v = to_f(base[f]);
    if (ln.mean) {
        const size_t n = img_row0 + f / kC;
        const int c = (int)(f % kC);
        v = (v - ln.mean[n]) * ln.rstd[n] * ln.gamma[c] + ln.beta[c];
    }
    return v;
}

template <typename T>
__global__ void gate_pool_kernel(const T* __restrict__ x, const T* __restrict__ y, LnRef lx, LnRef ly, float* __restrict__ pooled,
                                 uint8_t* __restrict__ amax, int HW) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= HW) return;
    const int b = blockIdx.y, z = blockIdx.z;
    const T* base = (z == 0 ? x : y) + (size_t)b * HW * kC;
    const LnRef ln = z == 0 ? lx : ly;
    float s = 0.f, mx = -INFINITY;
    int am = 0;
#pragma unroll 8
    for (int k = 0; k < kC; ++k) {
        const float v = normed_at(base, ln, (size_t)b * HW, (size_t)k * HW + j);
        s += v;
        if (v > mx) { mx = v; am = k; }
    }
    pooled[((size_t)b * 4 + z * 2 + 0) * HW + j] = s * (1.0f / kC);
    pooled[((size_t)b * 4 + z * 2 + 1) * HW + j] = mx;
    amax[((size_t)b * 2 + z) * HW + j] = (uint8_t)am;
}

// 8 consecutive flat positions of one "channel" row k per thread (HW % 8 == 0): one 16/32-byte load instead of eight scalar ones,
// and the 8 elements f0..f0+7 (f0 % 8 == 0) belong to ONE token, so the LayerNorm statistics are fetched once per vector.
// Same per-element expressions and the same k order as the scalar kernel above.
template <typename T>
__device__ __forceinline__ void normed8_at(const T* __restrict__ base, const LnRef& ln, size_t img_row0, size_t f0, float v[8]) {
    load8(base + f0, v);
    if (ln.mean) {
        const size_t n = img_row0 + f0 / kC;
        const int c0 = (int)(f0 % kC);
        const float mu = ln.mean[n], rs = ln.rstd[n];
        float g[8], b[8];
        load8(ln.gamma + c0, g);
        load8(ln.beta + c0, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (v[i] - mu) * rs * g[i] + b[i];
    }
}

// Four lanes share one group of 8 positions and split the 32 "channel" rows k between them (k = kq, kq + 4, ...: 8 rows each,
// all 8 loads in flight at once), then combine with two xor-shuffles -- 4x the loads in flight of a one-thread-per-group loop,
// which on this latency-bound pass (33.5 MB, 2 x 16 images) is what sets the run time.
template <typename T>
__global__ void __launch_bounds__(128) gate_pool_vec_kernel(const T* __restrict__ x, const T* __restrict__ y, LnRef lx, LnRef ly,
                                                            float* __restrict__ pooled, uint8_t* __restrict__ amax, int HW) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int kq = t & 3;
    int j0 = (t >> 2) * 8;
    const bool live = j0 < HW;
    if (!live) j0 = 0;                                   // keep the whole warp in the shuffles
    const int b = blockIdx.y, z = blockIdx.z;
    const T* base = (z == 0 ? x : y) + (size_t)b * HW * kC;
    const LnRef ln = z == 0 ? lx : ly;
    float s[8], mx[8];
    int am[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = 0.f; mx[i] = -INFINITY; am[i] = 0; }
#pragma unroll
    for (int kk = 0; kk < kC / 4; ++kk) {
        const int k = kk * 4 + kq;
        float v[8];
        normed8_at(base, ln, (size_t)b * HW, (size_t)k * HW + j0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s[i] += v[i];
            if (v[i] > mx[i]) { mx[i] = v[i]; am[i] = k; }          // k ascending within a lane: first occurrence wins
        }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
            const float om = __shfl_xor_sync(0xffffffffu, mx[i], o);
            const int oa = __shfl_xor_sync(0xffffffffu, am[i], o);
            if (om > mx[i] || (om == mx[i] && oa < am[i])) { mx[i] = om; am[i] = oa; }     // torch.max: first arg-max on ties
        }
    }
    if (!live || kq != 0) return;
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] *= (1.0f / kC);
    store8(pooled + ((size_t)b * 4 + z * 2 + 0) * HW + j0, s);
    store8(pooled + ((size_t)b * 4 + z * 2 + 1) * HW + j0, mx);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { lo |= (uint32_t)am[i] << (8 * i); hi |= (uint32_t)am[4 + i] << (8 * i); }
    *reinterpret_cast<uint2*>(amax + ((size_t)b * 2 + z) * HW + j0) = make_uint2(lo, hi);
}

// 32x8-pixel tile per block: the four pooled maps of the tile (+3-pixel halo, zero outside the image = the conv's padding) are
// staged in shared memory once, so the 2 x 2 x 49 taps per pixel are shared-memory reads instead of bounds-checked global loads.
constexpr int kGmTW = 32, kGmTH = 8, kGmHW = kGmTW + 6, kGmHH = kGmTH + 6;
__global__ void __launch_bounds__(kGmTW * kGmTH)
gate_map_kernel(const float* __restrict__ pooled, const float* __restrict__ w_sa1, const float* __restrict__ w_sa2,
                const float* __restrict__ lvl_w, const float* __restrict__ lvl_b,
                float* __restrict__ smap, float* __restrict__ gmap, int H, int W) {
    __shared__ float ws[2][98];
    __shared__ float tile[4][kGmHH][kGmHW + 1];
    for (int i = threadIdx.x; i < 196; i += blockDim.x) ws[i / 98][i % 98] = (i < 98 ? w_sa1[i] : w_sa2[i - 98]);
    const int HW = H * W, b = blockIdx.y;
    const int tiles_x = (W + kGmTW - 1) / kGmTW;
    const int h0 = (blockIdx.x / tiles_x) * kGmTH, w0 = (blockIdx.x % tiles_x) * kGmTW;
    for (int i = threadIdx.x; i < 4 * kGmHH * kGmHW; i += blockDim.x) {
        const int m = i / (kGmHH * kGmHW), r = (i / kGmHW) % kGmHH, c = i % kGmHW;
        const int yy = h0 + r - 3, xx = w0 + c - 3;
        tile[m][r][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(pooled + ((size_t)b * 4 + m) * HW + yy * W + xx) : 0.f;
    }
    __syncthreads();
    const int tr = threadIdx.x / kGmTW, tc = threadIdx.x % kGmTW;
    const int h = h0 + tr, w = w0 + tc;
    if (h >= H || w >= W) return;
    const int pix = h * W + w;
    float s[2];
#pragma unroll
    for (int z = 0; z < 2; ++z) {
        float acc = 0.f;
#pragma unroll
        for (int ci = 0; ci < 2; ++ci)
#pragma unroll
            for (int dy = 0; dy < 7; ++dy)                       // same tap order as the reference conv; out-of-image taps add 0
#pragma unroll
                for (int dx = 0; dx < 7; ++dx) acc += tile[z * 2 + ci][tr + dy][tc + dx] * ws[z][ci * 49 + dy * 7 + dx];
        s[z] = 1.0f / (1.0f + expf(-acc));
    }
    const float l0 = lvl_w[0] * s[0] + lvl_w[1] * s[1] + lvl_b[0];
    const float l1 = lvl_w[2] * s[0] + lvl_w[3] * s[1] + lvl_b[1];
    const float g0 = 1.0f / (1.0f + expf(l1 - l0));
    smap[((size_t)b * 2 + 0) * HW + pix] = s[0];
    smap[((size_t)b * 2 + 1) * HW + pix] = s[1];
    gmap[((size_t)b * 2 + 0) * HW + pix] = g0;
    gmap[((size_t)b * 2 + 1) * HW + pix] = 1.0f - g0;
}

// ------------------------------------------------------------------------------------------
// window helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int token_pixel(const WinGeom& g, int wi, int wj, int t) {
    const int hy = wi * kWS + t / kWS - g.ph, wx = wj * kWS + t % kWS - g.pw;
    return (hy >= 0 && hy < g.H && wx >= 0 && wx < g.W) ? hy * g.W + wx : -1;
}

// gated tokens of one window -> dst[49][36]; pad tokens are zeros
template <typename T>
__device__ __forceinline__ void load_window(const T* __restrict__ src, const LnRef& ln, const float* __restrict__ gate_b, float* dst,
                                            const WinGeom& g, int b, int wi, int wj) {
    for (int idx = threadIdx.x; idx < kL * 4; idx += kThreads) {
        const int t = idx >> 2, part = idx & 3;
        const int n = token_pixel(g, wi, wj, t);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (n >= 0) {
            load8(src + ((size_t)b * g.HW + n) * kC + part * 8, v);
            if (ln.mean) {
                const float mu = ln.mean[(size_t)b * g.HW + n], rs = ln.rstd[(size_t)b * g.HW + n];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = (v[i] - mu) * rs * ln.gamma[part * 8 + i] + ln.beta[part * 8 + i];
            }
            if (gate_b) {
                int gi = (int)(((uint32_t)n * kC + part * 8) % (uint32_t)g.HW);      // n < HW <= 2^26: 32-bit modulo
#pragma unroll
                for (int i = 0; i < 8; ++i) { v[i] *= gate_b[gi]; if (++gi == g.HW) gi = 0; }
            }
        }
        float4* d = reinterpret_cast<float4*>(dst + t * kLD + part * 8);
        d[0] = make_float4(v[0], v[1], v[2], v[3]);
        d[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

// dst[t][lane] = (sum_i src[t][i] * Wm[lane][i] + bias) * scale   for the 7 tokens of group tg
__device__ __forceinline__ void proj7(const float* src, const float* Wm, float bias, float scale, float* dst, int tg, int lane) {
    float acc[7];
#pragma unroll
    for (int tt = 0; tt < 7; ++tt) acc[tt] = bias;
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
        const float4 w = *reinterpret_cast<const float4*>(Wm + lane * kLD + i4 * 4);
#pragma unroll
        for (int tt = 0; tt < 7; ++tt) {
            const float4 xv = *reinterpret_cast<const float4*>(src + (tg * 7 + tt) * kLD + i4 * 4);
            acc[tt] += xv.x * w.x + xv.y * w.y + xv.z * w.z + xv.w * w.w;
        }
    }
#pragma unroll
    for (int tt = 0; tt < 7; ++tt) dst[(tg * 7 + tt) * kLD + lane] = acc[tt] * scale;
}

// acc[tt] += sum_c G[t][c] * Wm[c][lane]
__device__ __forceinline__ void projT7(const float* G, const float* Wm, float acc[7], int tg, int lane) {
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
        const float w0 = Wm[(c4 * 4 + 0) * kLD + lane], w1 = Wm[(c4 * 4 + 1) * kLD + lane];
        const float w2 = Wm[(c4 * 4 + 2) * kLD + lane], w3 = Wm[(c4 * 4 + 3) * kLD + lane];
#pragma unroll
        for (int tt = 0; tt < 7; ++tt) {
            const float4 gv = *reinterpret_cast<const float4*>(G + (tg * 7 + tt) * kLD + c4 * 4);
            acc[tt] += gv.x * w0 + gv.y * w1 + gv.z * w2 + gv.w * w3;
        }
    }
}

__device__ __forceinline__ void load_weights(float* Wsm, float* bsm, const rss_attn_params& p) {
    const float* wsrc[4] = {p.q_w, p.k_w, p.v_w, p.o_w};
    const float* bsrc[4] = {p.q_b, p.k_b, p.v_b, p.o_b};
    for (int idx = threadIdx.x; idx < 4 * kC * kC; idx += kThreads) {
        const int m = idx / (kC * kC), r = (idx / kC) % kC, c = idx % kC;
        Wsm[(m * kC + r) * kLD + c] = wsrc[m][r * kC + c];
    }
    for (int idx = threadIdx.x; idx < 4 * kC; idx += kThreads) bsm[idx] = bsrc[idx / kC][idx % kC];
}

__device__ __forceinline__ void load16(const float* p, float v[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 f = reinterpret_cast<const float4*>(p)[i];
        v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
    }
}

// channel gate of DAL.py:1003-1010: S2 = q_h^T k_h (16x16), gate = sigmoid(mean(S2) + max(S2)).
// Also returns the arg-max (first occurrence, row-major a*16+b) for the backward.
__device__ __forceinline__ void channel_gate(const float* q, const float* k, float* red /*>=16 floats*/,
                                             float* gate /*[2]*/, int* amax_idx /*[2]*/) {
    const int tid = threadIdx.x, h = tid >> 6, a = (tid & 63) >> 2, b0 = (tid & 3) * 4;
    float s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 7
    for (int t = 0; t < kL; ++t) {
        const float qa = q[t * kLD + h * kHD + a];
        const float4 kb = *reinterpret_cast<const float4*>(k + t * kLD + h * kHD + b0);
        s2[0] += qa * kb.x; s2[1] += qa * kb.y; s2[2] += qa * kb.z; s2[3] += qa * kb.w;
    }
    float sum = s2[0] + s2[1] + s2[2] + s2[3];
    float mx = s2[0];
    int mi = a * 16 + b0;
#pragma unroll
    for (int e = 1; e < 4; ++e) if (s2[e] > mx) { mx = s2[e]; mi = a * 16 + b0 + e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (om > mx || (om == mx && oi < mi)) { mx = om; mi = oi; }
    }
    const int warp = tid >> 5;
    if ((tid & 31) == 0) { red[warp] = sum; red[4 + warp] = mx; reinterpret_cast<int*>(red)[8 + warp] = mi; }
    __syncthreads();
    if (tid < 2) {
        const float s = red[2 * tid] + red[2 * tid + 1];
        float m0 = red[4 + 2 * tid], m1 = red[4 + 2 * tid + 1];
        int i0 = reinterpret_cast<int*>(red)[8 + 2 * tid], i1 = reinterpret_cast<int*>(red)[8 + 2 * tid + 1];
        if (m1 > m0 || (m1 == m0 && i1 < i0)) { m0 = m1; i0 = i1; }
        gate[tid] = 1.0f / (1.0f + expf(-(s * (1.0f / 256.0f) + m0)));
        amax_idx[tid] = i0;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
constexpr int kFwdSmemFloats = 5 * kTok + 4 * kC * kLD + 4 * kC + 32;

template <typename T>
__global__ void __launch_bounds__(kThreads, 4)
win_attn_fwd_kernel(const T* __restrict__ x, const T* __restrict__ y, LnRef lx, LnRef ly, const float* __restrict__ gmap,
                    const T* __restrict__ xres, T* __restrict__ out, rss_attn_params p, WinGeom g) {
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;
    float* ys = xs + kTok;
    float* q = ys + kTok;
    float* k = q + kTok;
    float* v = k + kTok;
    float* Wsm = v + kTok;
    float* bsm = Wsm + 4 * kC * kLD;
    float* misc = bsm + 4 * kC;            // [0..15] reduction scratch, [16..17] gate, [18..19] argmax
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    load_weights(Wsm, bsm, p);

    for (int win = blockIdx.x; win < g.nWin; win += gridDim.x) {
        const int b = win / (g.qh * g.qw), wi = (win / g.qw) % g.qh, wj = win % g.qw;
        load_window(x, lx, gmap + ((size_t)b * 2 + 0) * g.HW, xs, g, b, wi, wj);
        load_window(y, ly, gmap + ((size_t)b * 2 + 1) * g.HW, ys, g, b, wi, wj);
        __syncthreads();
        for (int task = warp; task < 21; task += 4) {      // q,k,v projections (DAL.py:873-875)
            const int m = task / 7, tg = task % 7;
            proj7(m == 0 ? xs : ys, Wsm + m * kC * kLD, bsm[m * kC + lane], m == 0 ? 0.25f : 1.0f,
                  m == 0 ? q : (m == 1 ? k : v), tg, lane);
        }
        __syncthreads();
        channel_gate(q, k, misc, misc + 16, reinterpret_cast<int*>(misc + 18));
        {   // one attention row per thread: S = q_i k^T, softmax, O_i = P_i v (DAL.py:959,996,1012)
            const int h = tid >> 6, i = tid & 63;
            if (i < kL) {
                float qi[16], S[kL], O[16];
                load16(q + i * kLD + h * kHD, qi);
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < kL; ++j) {
                    float kj[16];
                    load16(k + j * kLD + h * kHD, kj);
                    float s = 0.f;
#pragma unroll
                    for (int d = 0; d < 16; ++d) s += qi[d] * kj[d];
                    S[j] = s;
                    mx = fmaxf(mx, s);
                }
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < kL; ++j) { S[j] = __expf(S[j] - mx); sum += S[j]; }
#pragma unroll
                for (int d = 0; d < 16; ++d) O[d] = 0.f;
#pragma unroll
                for (int j = 0; j < kL; ++j) {
                    float vj[16];
                    load16(v + j * kLD + h * kHD, vj);
#pragma unroll
                    for (int d = 0; d < 16; ++d) O[d] += S[j] * vj[d];
                }
                const float sc = misc[16 + h] / sum;          // channel gate folded into the normaliser (DAL:1013)
#pragma unroll
                for (int d = 0; d < 16; ++d) xs[i * kLD + h * kHD + d] = O[d] * sc;   // xs is dead: reuse as merged O
            }
        }
        __syncthreads();
        for (int tg = warp; tg < 7; tg += 4)                 // out_proj (DAL.py:1020) -> ys (dead) as staging
            proj7(xs, Wsm + 3 * kC * kLD, bsm[3 * kC + lane], 1.0f, ys, tg, lane);
        __syncthreads();
        for (int idx = tid; idx < kL * 4; idx += kThreads) {   // un-window, crop, + residual (isa:384-390,415-426; MTFM:107)
            const int t = idx >> 2, part = idx & 3;
            const int n = token_pixel(g, wi, wj, t);
            if (n < 0) continue;
            const size_t off = ((size_t)b * g.HW + n) * kC + part * 8;
            float r[8];
            if (xres) load8(xres + off, r);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] += ys[t * kLD + part * 8 + i];
            store8(out + off, r);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// forward, tensor-core version for bf16 activations: the window's tokens, q/k/v and the weights sit in shared memory as
// bf16 (row stride 40 elements = 80 B: conflict-free ldmatrix), every product runs on mma.sync m16n8k16 (bf16 x bf16 -> fp32),
// the softmax stays in the accumulator registers and P feeds the PV product straight from registers.  49 tokens are padded to
// 64 rows (4 warps x 16 query rows); padded keys are masked out of the softmax.  ~10x fewer issued instructions than the
// fp32 SIMT kernel above (which remains the strict-parity path for fp32 activations).
// (49x16 tiles with K=16 cannot fill a 128-row tcgen05 tile: a UMMA per (window, head) would be >60 % padding and would need
// a TMEM round trip per 49x49 score tile; warp-level MMA keeps the whole window in registers instead.  See DESIGN.md.)
// ------------------------------------------------------------------------------------------
constexpr int kTS = 40;                       // bf16 row stride of the token / weight tiles
constexpr int kRows = 64;

__device__ __forceinline__ void mma_bf16(float d[4], const uint32_t a[4], const uint32_t b[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldsm_x4(uint32_t r[4], const __nv_bfloat16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t r[4], const __nv_bfloat16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2(uint32_t r[2], const __nv_bfloat16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t r[2], const __nv_bfloat16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// gated (and normalised) tokens of one window -> bf16 dst[64][40]; pad tokens and rows 49..63 are zeros
template <typename T>
__device__ __forceinline__ void load_window_bf16(const T* __restrict__ src, const LnRef& ln, const float* __restrict__ gate_b,
                                                 __nv_bfloat16* dst, const WinGeom& g, int b, int wi, int wj) {
    // idx = threadIdx.x + k * kThreads and kThreads % 4 == 0: a thread always handles the same 8-channel part, so its LayerNorm
    // weights are loaded once per window instead of once per token
    const int part = threadIdx.x & 3;
    float gm[8], bt[8];
    if (ln.mean) {
        load8(ln.gamma + part * 8, gm);
        load8(ln.beta + part * 8, bt);
    }
    const bool gate_vec = (g.HW & 7) == 0;                   // then the 8 gate values of a part never wrap around HW
    for (int idx = threadIdx.x; idx < kRows * 4; idx += kThreads) {
        const int t = idx >> 2;
        const int n = t < kL ? token_pixel(g, wi, wj, t) : -1;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (n >= 0) {
            load8(src + ((size_t)b * g.HW + n) * kC + part * 8, v);
            if (ln.mean) {
                const float mu = ln.mean[(size_t)b * g.HW + n], rs = ln.rstd[(size_t)b * g.HW + n];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = (v[i] - mu) * rs * gm[i] + bt[i];
            }
            if (gate_b) {
                int gi = (int)(((uint32_t)n * kC + part * 8) % (uint32_t)g.HW);      // n < HW <= 2^26: 32-bit modulo
                if (gate_vec) {
                    float gv[8];
                    load8(gate_b + gi, gv);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] *= gv[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { v[i] *= gate_b[gi]; if (++gi == g.HW) gi = 0; }
                }
            }
        }
        *reinterpret_cast<uint4*>(dst + t * kTS + part * 8) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
}

// L2 prefetch of the token rows of the window this CTA processes NEXT (persistent loop, stride gridDim.x): the loads of a window sit at
// the head of a long dependent chain with nothing of the same CTA to overlap them (13 % of the forward kernel's stall samples were
// the first use of the loaded tokens); issued right after the current window's loads, the lines are in L2 by the time they are needed.
template <typename T>
__device__ __forceinline__ void prefetch_window_l2(const T* __restrict__ a, const T* __restrict__ b2, const T* __restrict__ c,
                                                   const WinGeom& g, int win) {
    if (win >= g.nWin || threadIdx.x >= kL) return;
    const int b = win / (g.qh * g.qw), wi = (win / g.qw) % g.qh, wj = win % g.qw;
    const int n = token_pixel(g, wi, wj, threadIdx.x);
    if (n < 0) return;
    const size_t off = ((size_t)b * g.HW + n) * kC;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a + off));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + off));
    if (c) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
}

// The four 32x32 projection matrices -> bf16 [m][row][kTS] in shared memory, biases -> bsm.  Eight 16-byte loads per thread, all in
// flight before the first conversion: the scalar version (one dependent 4-byte load per element, 32 per thread) was 12.6 % of the
// forward kernel's stall samples -- a CTA only amortises this prologue over ~6.5 windows (profiles/ncu_r2_attn_fwd_source_lines.txt).
__device__ __forceinline__ void stage_proj_weights_bf16(const rss_attn_params& p, __nv_bfloat16* Wsm, float* bsm, int tid) {
    static_assert(kThreads == 128 && kC == 32, "index math below");
    const float* w0 = p.q_w; const float* w1 = p.k_w; const float* w2 = p.v_w; const float* w3 = p.o_w;
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {                           // float4 index tid + 128 i: matrix i >> 1, element (tid + 128 (i & 1)) * 4
        const float* w = (i >> 1) == 0 ? w0 : ((i >> 1) == 1 ? w1 : ((i >> 1) == 2 ? w2 : w3));
        v[i] = __ldg(reinterpret_cast<const float4*>(w) + tid + 128 * (i & 1));
    }
    const float* bs = (tid >> 5) == 0 ? p.q_b : ((tid >> 5) == 1 ? p.k_b : ((tid >> 5) == 2 ? p.v_b : p.o_b));
    const float bv = bs[tid & 31];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rem = tid + 128 * (i & 1), r = rem >> 3, c = (rem & 7) * 4;
        *reinterpret_cast<uint2*>(Wsm + ((i >> 1) * kC + r) * kTS + c) = make_uint2(pack_bf16(v[i].x, v[i].y), pack_bf16(v[i].z, v[i].w));
    }
    bsm[tid] = bv;
}

// acc[nt] (16 rows x 8 cols each) = A[16 rows of this warp][32] . W[n][k]^T for the 4 n-tiles of a 32-wide projection
__device__ __forceinline__ void proj_mma(const __nv_bfloat16* A, const __nv_bfloat16* Wm, int row0, int lane, float acc[4][4]) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        uint32_t a[4];
        ldsm_x4(a, A + (row0 + (lane & 15)) * kTS + ks * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            uint32_t bfr[2];
            ldsm_x2(bfr, Wm + (nt * 8 + (lane & 7)) * kTS + ks * 16 + ((lane >> 3) & 1) * 8);
            mma_bf16(acc[nt], a, bfr);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 6)
win_attn_fwd_tc_kernel(const T* __restrict__ x, const T* __restrict__ y, LnRef lx, LnRef ly, const float* __restrict__ gmap,
                       const T* __restrict__ xres, T* __restrict__ out, rss_attn_params p, WinGeom g) {
    __shared__ __align__(16) __nv_bfloat16 xs[kRows * kTS], ys[kRows * kTS], qkv[3 * kRows * kTS], Wsm[4 * kC * kTS];
    __shared__ float bsm[4 * kC], gate[2];
    __nv_bfloat16* qs = qkv;
    __nv_bfloat16* ks = qkv + kRows * kTS;
    __nv_bfloat16* vs = qkv + 2 * kRows * kTS;
    float* stage = reinterpret_cast<float*>(qkv);          // [64][36] fp32 output staging, aliases q and k once they are dead
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
    stage_proj_weights_bf16(p, Wsm, bsm, tid);
    const int row0 = warp * 16;

    for (int win = blockIdx.x; win < g.nWin; win += gridDim.x) {
        const int b = win / (g.qh * g.qw), wi = (win / g.qw) % g.qh, wj = win % g.qw;
        load_window_bf16(x, lx, gmap + ((size_t)b * 2 + 0) * g.HW, xs, g, b, wi, wj);
        load_window_bf16(y, ly, gmap + ((size_t)b * 2 + 1) * g.HW, ys, g, b, wi, wj);
        prefetch_window_l2(x, y, (const T*)nullptr, g, win + (int)gridDim.x);
        __syncthreads();
        // ---- q, k, v projections (DAL.py:873-875); rows >= 49 are MMA padding and are stored as zeros
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            float acc[4][4];
            proj_mma(m == 0 ? xs : ys, Wsm + m * kC * kTS, row0, lane, acc);
            __nv_bfloat16* dst = m == 0 ? qs : (m == 1 ? ks : vs);
            const float sc = m == 0 ? 0.25f : 1.0f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int col = nt * 8 + tq * 2;
                const float b0 = bsm[m * kC + col], b1 = bsm[m * kC + col + 1];
                const int r0 = row0 + gq, r1 = r0 + 8;
                *reinterpret_cast<uint32_t*>(dst + r0 * kTS + col) = r0 < kL ? pack_bf16((acc[nt][0] + b0) * sc, (acc[nt][1] + b1) * sc) : 0u;
                *reinterpret_cast<uint32_t*>(dst + r1 * kTS + col) = r1 < kL ? pack_bf16((acc[nt][2] + b0) * sc, (acc[nt][3] + b1) * sc) : 0u;
            }
        }
        __syncthreads();
        // ---- channel gate (DAL.py:1003-1010): S2 = q_h^T k_h (16x16) by warp h, gate = sigmoid(mean + max)
        if (warp < 2) {
            const int h = warp;
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int kt = 0; kt < 4; ++kt) {
                uint32_t a[4];
                const int i = lane >> 3;
                ldsm_x4_t(a, qs + (kt * 16 + (lane & 7) + 8 * (i >> 1)) * kTS + h * kHD + 8 * (i & 1));
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) {
                    uint32_t bfr[2];
                    ldsm_x2_t(bfr, ks + (kt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kTS + h * kHD + nb * 8);
                    mma_bf16(acc[nb], a, bfr);
                }
            }
            float sum = 0.f, mx = -INFINITY;
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int e = 0; e < 4; ++e) { sum += acc[nb][e]; mx = fmaxf(mx, acc[nb][e]); }
            sum = warp_sum(sum);
            mx = warp_max(mx);
            if (lane == 0) gate[h] = 1.0f / (1.0f + expf(-(sum * (1.0f / 256.0f) + mx)));
        }
        __syncthreads();
        // ---- attention rows of this warp, both heads: S = q k^T, softmax, O = P v, gate folded into the normaliser
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t aq[4];
            ldsm_x4(aq, qs + (row0 + (lane & 15)) * kTS + h * kHD + (lane >> 4) * 8);
            float S[7][4];
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                S[j][0] = S[j][1] = S[j][2] = S[j][3] = 0.f;
                uint32_t bk[2];
                ldsm_x2(bk, ks + (j * 8 + (lane & 7)) * kTS + h * kHD + ((lane >> 3) & 1) * 8);
                mma_bf16(S[j], aq, bk);
            }
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const int c = j * 8 + tq * 2;
                if (c >= kL) { S[j][0] = -INFINITY; S[j][2] = -INFINITY; }
                if (c + 1 >= kL) { S[j][1] = -INFINITY; S[j][3] = -INFINITY; }
                m0 = fmaxf(m0, fmaxf(S[j][0], S[j][1]));
                m1 = fmaxf(m1, fmaxf(S[j][2], S[j][3]));
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                S[j][0] = __expf(S[j][0] - m0); S[j][1] = __expf(S[j][1] - m0);
                S[j][2] = __expf(S[j][2] - m1); S[j][3] = __expf(S[j][3] - m1);
                s0 += S[j][0] + S[j][1];
                s1 += S[j][2] + S[j][3];
            }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            float O[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {                 // keys 16kk..16kk+15: P straight from the score registers
                uint32_t ap[4];
                ap[0] = pack_bf16(S[2 * kk][0], S[2 * kk][1]);
                ap[1] = pack_bf16(S[2 * kk][2], S[2 * kk][3]);
                ap[2] = kk < 3 ? pack_bf16(S[2 * kk + 1 < 7 ? 2 * kk + 1 : 6][0], S[2 * kk + 1 < 7 ? 2 * kk + 1 : 6][1]) : 0u;
                ap[3] = kk < 3 ? pack_bf16(S[2 * kk + 1 < 7 ? 2 * kk + 1 : 6][2], S[2 * kk + 1 < 7 ? 2 * kk + 1 : 6][3]) : 0u;
#pragma unroll
                for (int nd = 0; nd < 2; ++nd) {
                    uint32_t bv[2];
                    ldsm_x2_t(bv, vs + (kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kTS + h * kHD + nd * 8);
                    mma_bf16(O[nd], ap, bv);
                }
            }
            const float sc0 = gate[h] / s0, sc1 = gate[h] / s1;
#pragma unroll
            for (int nd = 0; nd < 2; ++nd) {
                const int col = h * kHD + nd * 8 + tq * 2;
                *reinterpret_cast<uint32_t*>(xs + (row0 + gq) * kTS + col) = pack_bf16(O[nd][0] * sc0, O[nd][1] * sc0);
                *reinterpret_cast<uint32_t*>(xs + (row0 + gq + 8) * kTS + col) = pack_bf16(O[nd][2] * sc1, O[nd][3] * sc1);
            }
        }
        __syncthreads();                                     // q/k/v dead everywhere: their storage becomes the fp32 staging tile
        {
            float acc[4][4];
            proj_mma(xs, Wsm + 3 * kC * kTS, row0, lane, acc);   // out_proj (DAL.py:1020); xs now holds the merged heads
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int col = nt * 8 + tq * 2;
                const float b0 = bsm[3 * kC + col], b1 = bsm[3 * kC + col + 1];
                stage[(row0 + gq) * kLD + col] = acc[nt][0] + b0; stage[(row0 + gq) * kLD + col + 1] = acc[nt][1] + b1;
                stage[(row0 + gq + 8) * kLD + col] = acc[nt][2] + b0; stage[(row0 + gq + 8) * kLD + col + 1] = acc[nt][3] + b1;
            }
        }
        __syncthreads();
        for (int idx = tid; idx < kL * 4; idx += kThreads) {   // un-window, crop, + residual
            const int t = idx >> 2, part = idx & 3;
            const int n = token_pixel(g, wi, wj, t);
            if (n < 0) continue;
            const size_t off = ((size_t)b * g.HW + n) * kC + part * 8;
            float r[8];
            if (xres) load8(xres + off, r);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] += stage[t * kLD + part * 8 + i];
            store8(out + off, r);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// backward of the windowed Mhca: produces d(gated x tokens), d(gated y tokens), dW/db of the 4 projections
// ------------------------------------------------------------------------------------------
constexpr int kBwdSmemFloats = 7 * kTok + 2 * 2 * kL * kPS + 4 * kC * kLD + 4 * kC + 32;

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
win_attn_bwd_kernel(const T* __restrict__ x, const T* __restrict__ y, LnRef lx, LnRef ly, const float* __restrict__ gmap,
                    const T* __restrict__ dout, float* __restrict__ dxg, float* __restrict__ dyg,
                    rss_attn_params p, rss_attn_grads gr, WinGeom g) {
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;
    float* ys = xs + kTok;
    float* q = ys + kTok;
    float* k = q + kTok;
    float* v = k + kTok;
    float* Om = v + kTok;
    float* dOm = Om + kTok;
    float* Pb = dOm + kTok;                 // [2][49][49]   (first used to stage the dout tile)
    float* dSb = Pb + 2 * kL * kPS;         // [2][49][49]
    float* Wsm = dSb + 2 * kL * kPS;
    float* bsm = Wsm + 4 * kC * kLD;
    float* misc = bsm + 4 * kC;             // [0..15] scratch, [16..17] gate, [18..19] argmax, [20..23] dgate partials
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    load_weights(Wsm, bsm, p);

    // persistent per-thread weight-gradient accumulators: matrix m = warp, output channel c = lane
    float accW[kC], accB = 0.f;
#pragma unroll
    for (int i = 0; i < kC; ++i) accW[i] = 0.f;

    for (int win = blockIdx.x; win < g.nWin; win += gridDim.x) {
        const int b = win / (g.qh * g.qw), wi = (win / g.qw) % g.qh, wj = win % g.qw;
        load_window(x, lx, gmap + ((size_t)b * 2 + 0) * g.HW, xs, g, b, wi, wj);
        load_window(y, ly, gmap + ((size_t)b * 2 + 1) * g.HW, ys, g, b, wi, wj);
        load_window(dout, LnRef{nullptr, nullptr, nullptr, nullptr}, (const float*)nullptr, Pb, g, b, wi, wj);   // cropped grad: pad queries get 0
        __syncthreads();
        for (int task = warp; task < 28; task += 4) {
            const int m = task / 7, tg = task % 7;
            if (m < 3) {
                proj7(m == 0 ? xs : ys, Wsm + m * kC * kLD, bsm[m * kC + lane], m == 0 ? 0.25f : 1.0f,
                      m == 0 ? q : (m == 1 ? k : v), tg, lane);
            } else {                                                    // dOm = dout . Wo
                float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                projT7(Pb, Wsm + 3 * kC * kLD, acc, tg, lane);
#pragma unroll
                for (int tt = 0; tt < 7; ++tt) dOm[(tg * 7 + tt) * kLD + lane] = acc[tt];
            }
        }
        __syncthreads();
        channel_gate(q, k, misc, misc + 16, reinterpret_cast<int*>(misc + 18));

        const int h = tid >> 6, i = tid & 63;
        const bool row_live = i < kL;
        float dq[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) dq[d] = 0.f;
        float dgate_part = 0.f;
        if (row_live) {
            // scores/probabilities of this thread's row live in Pb (own row only, stride 49: conflict-free)
            float qi[16], A[16], dA[16];
            float* Prow = Pb + (h * kL + i) * kPS;
            float* dSrow = dSb + (h * kL + i) * kPS;
            load16(q + i * kLD + h * kHD, qi);
            float mx = -INFINITY;
#pragma unroll 7
            for (int j = 0; j < kL; ++j) {
                float kj[16];
                load16(k + j * kLD + h * kHD, kj);
                float s = 0.f;
#pragma unroll
                for (int d = 0; d < 16; ++d) s += qi[d] * kj[d];
                Prow[j] = s;
                mx = fmaxf(mx, s);
            }
            float sum = 0.f;
#pragma unroll
            for (int d = 0; d < 16; ++d) A[d] = 0.f;
#pragma unroll 7
            for (int j = 0; j < kL; ++j) {
                const float e = __expf(Prow[j] - mx);
                sum += e;
                Prow[j] = e;
                float vj[16];
                load16(v + j * kLD + h * kHD, vj);
#pragma unroll
                for (int d = 0; d < 16; ++d) A[d] += e * vj[d];
            }
            const float inv = 1.0f / sum;
            const float gate = misc[16 + h];
            float rowdot = 0.f;
            load16(dOm + i * kLD + h * kHD, dA);
#pragma unroll
            for (int d = 0; d < 16; ++d) {
                A[d] *= inv;                                           // A = P v (un-gated)
                Om[i * kLD + h * kHD + d] = A[d] * gate;
                dgate_part += dA[d] * A[d];
                dA[d] *= gate;                                         // dA = dO * gate
                rowdot += dA[d] * A[d];                                // = sum_j P_ij dP_ij
                dOm[i * kLD + h * kHD + d] = dA[d];
            }
#pragma unroll 7
            for (int j = 0; j < kL; ++j) {
                float vj[16], kj[16];
                load16(v + j * kLD + h * kHD, vj);
                float dP = 0.f;
#pragma unroll
                for (int d = 0; d < 16; ++d) dP += dA[d] * vj[d];
                const float pij = Prow[j] * inv;
                const float dS = pij * (dP - rowdot);
                Prow[j] = pij;
                dSrow[j] = dS;
                load16(k + j * kLD + h * kHD, kj);
#pragma unroll
                for (int d = 0; d < 16; ++d) dq[d] += dS * kj[d];
            }
        }
        dgate_part = warp_sum(dgate_part);
        if (lane == 0) misc[20 + warp] = dgate_part;
        __syncthreads();
        {   // key/value side: thread owns key j=i of head h; plus the Q^T K gate terms (DAL.py:1003-1010)
            float dk[16], dv[16];
#pragma unroll
            for (int d = 0; d < 16; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
            if (row_live) {
                for (int r = 0; r < kL; ++r) {
                    const float ds = dSb[(h * kL + r) * kPS + i], pp = Pb[(h * kL + r) * kPS + i];
                    float qr[16], dAr[16];
                    load16(q + r * kLD + h * kHD, qr);
                    load16(dOm + r * kLD + h * kHD, dAr);
#pragma unroll
                    for (int d = 0; d < 16; ++d) { dk[d] += ds * qr[d]; dv[d] += pp * dAr[d]; }
                }
                const float gate = misc[16 + h];
                const float dz = (misc[20 + 2 * h] + misc[20 + 2 * h + 1]) * gate * (1.0f - gate);
                const int am = reinterpret_cast<int*>(misc + 18)[h], as = am >> 4, bs = am & 15;
                float qt[16], kt[16];
                load16(q + i * kLD + h * kHD, qt);
                load16(k + i * kLD + h * kHD, kt);
                float qsum = 0.f, ksum = 0.f, q_as = 0.f, k_bs = 0.f;
#pragma unroll
                for (int d = 0; d < 16; ++d) {
                    qsum += qt[d]; ksum += kt[d];
                    q_as = (d == as) ? qt[d] : q_as;
                    k_bs = (d == bs) ? kt[d] : k_bs;
                }
                const float u = dz * (1.0f / 256.0f);
#pragma unroll
                for (int d = 0; d < 16; ++d) {
                    dq[d] += u * ksum + ((d == as) ? dz * k_bs : 0.f);
                    dk[d] += u * qsum + ((d == bs) ? dz * q_as : 0.f);
                }
            }
            __syncthreads();                 // every read of q/k/v is done: recycle them as gradient tiles
            if (row_live) {
#pragma unroll
                for (int d = 0; d < 16; ++d) {
                    q[i * kLD + h * kHD + d] = 0.25f * dq[d];          // grad at q_proj output (scaling, DAL:873)
                    k[i * kLD + h * kHD + d] = dk[d];
                    v[i * kLD + h * kHD + d] = dv[d];
                }
            }
        }
        __syncthreads();
        for (int task = warp; task < 14; task += 4) {                   // grads w.r.t. the gated tokens
            const int side = task / 7, tg = task % 7;
            float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (side == 0) projT7(q, Wsm + 0 * kC * kLD, acc, tg, lane);
            else { projT7(k, Wsm + 1 * kC * kLD, acc, tg, lane); projT7(v, Wsm + 2 * kC * kLD, acc, tg, lane); }
            float* dst = side == 0 ? dxg : dyg;
#pragma unroll
            for (int tt = 0; tt < 7; ++tt) {
                const int n = token_pixel(g, wi, wj, tg * 7 + tt);
                if (n >= 0) dst[((size_t)b * g.HW + n) * kC + lane] = acc[tt];
            }
        }
        {   // weight / bias gradients: thread (m=warp, c=lane) accumulates dW_m[c][:] += sum_t G_m[t][c] * X_m[t][:]
            const float* G = warp == 0 ? q : (warp == 1 ? k : v);
            const float* X = warp == 0 ? xs : (warp == 3 ? Om : ys);
            for (int t = 0; t < kL; ++t) {
                float gv;
                if (warp < 3) gv = G[t * kLD + lane];
                else {
                    const int n = token_pixel(g, wi, wj, t);
                    gv = n >= 0 ? to_f(dout[((size_t)b * g.HW + n) * kC + lane]) : 0.f;
                }
                accB += gv;
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 xv = *reinterpret_cast<const float4*>(X + t * kLD + i4 * 4);
                    accW[i4 * 4 + 0] += gv * xv.x; accW[i4 * 4 + 1] += gv * xv.y;
                    accW[i4 * 4 + 2] += gv * xv.z; accW[i4 * 4 + 3] += gv * xv.w;
                }
            }
        }
        __syncthreads();
    }
    float* dW = warp == 0 ? gr.q_w : (warp == 1 ? gr.k_w : (warp == 2 ? gr.v_w : gr.o_w));
    float* dB = warp == 0 ? gr.q_b : (warp == 1 ? gr.k_b : (warp == 2 ? gr.v_b : gr.o_b));
#pragma unroll
    for (int i = 0; i < kC; ++i) atomicAdd(dW + lane * kC + i, accW[i]);
    atomicAdd(dB + lane, accB);
}

// ------------------------------------------------------------------------------------------
// backward, tensor-core version for bf16 activations (same conventions as win_attn_fwd_tc_kernel).  Per window:
//   recompute q,k,v, gate (+ arg-max), P;  dOm = dout.Wo;  A = P v;  dA = dOm*gate;  dP = dA v^T;  dS = P o (dP - rowdot);
//   dq = dS k;  dk = dS^T q;  dv = P^T dA  (P, dS, dA staged in smem as bf16 for the transposed products);
//   Q^T K gate terms added on the fp32 accumulators;  d(tokens) = G.W;  dW += G^T X  (persistent register accumulators).
// ------------------------------------------------------------------------------------------
constexpr int kPB = 72;                                  // bf16 row stride of the P / dS tiles (64 keys + pad, 144 B rows)
// 9 token tiles + P, dS of ONE head + 4 weight matrices: 75.4 KB -> 3 CTAs/SM (was 104 KB / 2 CTAs with both heads' P, dS resident)
constexpr int kBwdTcSmem = (9 * kRows * kTS + 2 * kRows * kPB + 4 * kC * kTS) * 2 + (4 * kC + 32) * 4;

// acc[nt] = A[16 rows][32] . W[k][n]  (W stored [k][n] row-major: the transposed use of a (out,in) weight)
__device__ __forceinline__ void projT_mma(const __nv_bfloat16* A, const __nv_bfloat16* Wm, int row0, int lane, float acc[4][4], bool zero) {
    if (zero) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        uint32_t a[4];
        ldsm_x4(a, A + (row0 + (lane & 15)) * kTS + ks * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            uint32_t bfr[2];
            ldsm_x2_t(bfr, Wm + (ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kTS + nt * 8);
            mma_bf16(acc[nt], a, bfr);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 3)
win_attn_bwd_tc_kernel(const T* __restrict__ x, const T* __restrict__ y, LnRef lx, LnRef ly, const float* __restrict__ gmap,
                       const T* __restrict__ dout, float* __restrict__ dxg, float* __restrict__ dyg,
                       rss_attn_params p, rss_attn_grads gr, WinGeom g) {
    extern __shared__ __align__(16) uint8_t smem_tc[];
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_tc);
    __nv_bfloat16* ys = xs + kRows * kTS;
    __nv_bfloat16* ds = ys + kRows * kTS;
    __nv_bfloat16* qs = ds + kRows * kTS;
    __nv_bfloat16* ks = qs + kRows * kTS;
    __nv_bfloat16* vs = ks + kRows * kTS;
    __nv_bfloat16* Om = vs + kRows * kTS;
    __nv_bfloat16* dAb = Om + kRows * kTS;
    __nv_bfloat16* Gq = dAb + kRows * kTS;
    // dk, dv overwrite k, v: head h's columns of k / v are dead once the query side of head h is done (block barrier below), the
    // key side reads only its own rows of k (gate terms) before it writes them, and nothing reads them again before the next window
    __nv_bfloat16* Gk = ks;
    __nv_bfloat16* Gv = vs;
    __nv_bfloat16* Pb = Gq + kRows * kTS;                 // [64][72], one head at a time
    __nv_bfloat16* dSb = Pb + kRows * kPB;                // [64][72]
    __nv_bfloat16* Wsm = dSb + kRows * kPB;               // [4][32][40]
    float* bsm = reinterpret_cast<float*>(Wsm + 4 * kC * kTS);
    float* misc = bsm + 4 * kC;                           // [0..1] gate, [2..3] argmax (int), [4..11] dgate partials [h][warp]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
    stage_proj_weights_bf16(p, Wsm, bsm, tid);
    const int row0 = warp * 16;
    // persistent weight-gradient accumulators: warp m owns matrix m; tile (mt, nt): rows c = mt*16 + gq (+8), cols i = nt*8 + 2tq (+1)
    float accW[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < 4; ++e) accW[a][b][e] = 0.f;
    float accB = 0.f;

    for (int win = blockIdx.x; win < g.nWin; win += gridDim.x) {
        const int b = win / (g.qh * g.qw), wi = (win / g.qw) % g.qh, wj = win % g.qw;
        load_window_bf16(x, lx, gmap + ((size_t)b * 2 + 0) * g.HW, xs, g, b, wi, wj);
        load_window_bf16(y, ly, gmap + ((size_t)b * 2 + 1) * g.HW, ys, g, b, wi, wj);
        load_window_bf16(dout, LnRef{nullptr, nullptr, nullptr, nullptr}, (const float*)nullptr, ds, g, b, wi, wj);
        prefetch_window_l2(x, y, dout, g, win + (int)gridDim.x);
        __syncthreads();
        // ---- recompute q, k, v ; dOm = dout . Wo (kept in registers)
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            float acc[4][4];
            proj_mma(m == 0 ? xs : ys, Wsm + m * kC * kTS, row0, lane, acc);
            __nv_bfloat16* dst = m == 0 ? qs : (m == 1 ? ks : vs);
            const float sc = m == 0 ? 0.25f : 1.0f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int col = nt * 8 + tq * 2;
                const float b0 = bsm[m * kC + col], b1 = bsm[m * kC + col + 1];
                const int r0 = row0 + gq, r1 = r0 + 8;
                *reinterpret_cast<uint32_t*>(dst + r0 * kTS + col) = r0 < kL ? pack_bf16((acc[nt][0] + b0) * sc, (acc[nt][1] + b1) * sc) : 0u;
                *reinterpret_cast<uint32_t*>(dst + r1 * kTS + col) = r1 < kL ? pack_bf16((acc[nt][2] + b0) * sc, (acc[nt][3] + b1) * sc) : 0u;
            }
        }
        float dOm[4][4];
        projT_mma(ds, Wsm + 3 * kC * kTS, row0, lane, dOm, true);
        __syncthreads();
        // ---- channel gate and its arg-max
        if (warp < 2) {
            const int h = warp;
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int kt = 0; kt < 4; ++kt) {
                uint32_t a[4];
                const int i = lane >> 3;
                ldsm_x4_t(a, qs + (kt * 16 + (lane & 7) + 8 * (i >> 1)) * kTS + h * kHD + 8 * (i & 1));
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) {
                    uint32_t bfr[2];
                    ldsm_x2_t(bfr, ks + (kt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kTS + h * kHD + nb * 8);
                    mma_bf16(acc[nb], a, bfr);
                }
            }
            float sum = 0.f, mx = -INFINITY;
            int mi = 0;
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int aa = gq + (e >> 1) * 8, bb = nb * 8 + tq * 2 + (e & 1), id = aa * 16 + bb;
                    sum += acc[nb][e];
                    if (acc[nb][e] > mx || (acc[nb][e] == mx && id < mi)) { mx = acc[nb][e]; mi = id; }
                }
            sum = warp_sum(sum);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float om = __shfl_xor_sync(0xffffffffu, mx, o);
                const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
                if (om > mx || (om == mx && oi < mi)) { mx = om; mi = oi; }
            }
            if (lane == 0) { misc[h] = 1.0f / (1.0f + expf(-(sum * (1.0f / 256.0f) + mx))); reinterpret_cast<int*>(misc)[2 + h] = mi; }
        }
        __syncthreads();
        // ---- per head: P, A = P v, dA, dP, dS, dq (rows of this warp)
        float dq[2][2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t aq[4];
            ldsm_x4(aq, qs + (row0 + (lane & 15)) * kTS + h * kHD + (lane >> 4) * 8);
            float S[7][4];
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                S[j][0] = S[j][1] = S[j][2] = S[j][3] = 0.f;
                uint32_t bk[2];
                ldsm_x2(bk, ks + (j * 8 + (lane & 7)) * kTS + h * kHD + ((lane >> 3) & 1) * 8);
                mma_bf16(S[j], aq, bk);
            }
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const int c = j * 8 + tq * 2;
                if (c >= kL) { S[j][0] = -INFINITY; S[j][2] = -INFINITY; }
                if (c + 1 >= kL) { S[j][1] = -INFINITY; S[j][3] = -INFINITY; }
                m0 = fmaxf(m0, fmaxf(S[j][0], S[j][1]));
                m1 = fmaxf(m1, fmaxf(S[j][2], S[j][3]));
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                S[j][0] = __expf(S[j][0] - m0); S[j][1] = __expf(S[j][1] - m0);
                S[j][2] = __expf(S[j][2] - m1); S[j][3] = __expf(S[j][3] - m1);
                s0 += S[j][0] + S[j][1];
                s1 += S[j][2] + S[j][3];
            }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            const float i0 = 1.0f / s0, i1 = 1.0f / s1;
#pragma unroll
            for (int j = 0; j < 7; ++j) { S[j][0] *= i0; S[j][1] *= i0; S[j][2] *= i1; S[j][3] *= i1; }   // P
            // P -> smem (bf16) for dv = P^T dA
            __nv_bfloat16* Prow = Pb + row0 * kPB;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                *reinterpret_cast<uint32_t*>(Prow + gq * kPB + j * 8 + tq * 2) = pack_bf16(S[j][0], S[j][1]);
                *reinterpret_cast<uint32_t*>(Prow + (gq + 8) * kPB + j * 8 + tq * 2) = pack_bf16(S[j][2], S[j][3]);
            }
            *reinterpret_cast<uint32_t*>(Prow + gq * kPB + 56 + tq * 2) = 0u;
            *reinterpret_cast<uint32_t*>(Prow + (gq + 8) * kPB + 56 + tq * 2) = 0u;
            // A = P v
            float A[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t ap[4];
                ap[0] = pack_bf16(S[2 * kk][0], S[2 * kk][1]);
                ap[1] = pack_bf16(S[2 * kk][2], S[2 * kk][3]);
                ap[2] = kk < 3 ? pack_bf16(S[kk < 3 ? 2 * kk + 1 : 6][0], S[kk < 3 ? 2 * kk + 1 : 6][1]) : 0u;
                ap[3] = kk < 3 ? pack_bf16(S[kk < 3 ? 2 * kk + 1 : 6][2], S[kk < 3 ? 2 * kk + 1 : 6][3]) : 0u;
#pragma unroll
                for (int nd = 0; nd < 2; ++nd) {
                    uint32_t bv[2];
                    ldsm_x2_t(bv, vs + (kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kTS + h * kHD + nd * 8);
                    mma_bf16(A[nd], ap, bv);
                }
            }
            const float gate = misc[h];
            float dA[2][4], dgp = 0.f, rd0 = 0.f, rd1 = 0.f;
#pragma unroll
            for (int nd = 0; nd < 2; ++nd) {
                const int col = h * kHD + nd * 8 + tq * 2;
                *reinterpret_cast<uint32_t*>(Om + (row0 + gq) * kTS + col) = pack_bf16(A[nd][0] * gate, A[nd][1] * gate);
                *reinterpret_cast<uint32_t*>(Om + (row0 + gq + 8) * kTS + col) = pack_bf16(A[nd][2] * gate, A[nd][3] * gate);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float d = dOm[2 * h + nd][e];
                    dgp += d * A[nd][e];
                    dA[nd][e] = d * gate;
                }
                rd0 += dA[nd][0] * A[nd][0] + dA[nd][1] * A[nd][1];
                rd1 += dA[nd][2] * A[nd][2] + dA[nd][3] * A[nd][3];
                *reinterpret_cast<uint32_t*>(dAb + (row0 + gq) * kTS + col) = pack_bf16(dA[nd][0], dA[nd][1]);
                *reinterpret_cast<uint32_t*>(dAb + (row0 + gq + 8) * kTS + col) = pack_bf16(dA[nd][2], dA[nd][3]);
            }
            rd0 += __shfl_xor_sync(0xffffffffu, rd0, 1); rd0 += __shfl_xor_sync(0xffffffffu, rd0, 2);
            rd1 += __shfl_xor_sync(0xffffffffu, rd1, 1); rd1 += __shfl_xor_sync(0xffffffffu, rd1, 2);
            dgp = warp_sum(dgp);
            if (lane == 0) misc[4 + h * 4 + warp] = dgp;
            // dP = dA v^T (K = 16 head dims): A operand straight from the dA accumulators
            uint32_t ada[4] = {pack_bf16(dA[0][0], dA[0][1]), pack_bf16(dA[0][2], dA[0][3]),
                               pack_bf16(dA[1][0], dA[1][1]), pack_bf16(dA[1][2], dA[1][3])};
            __nv_bfloat16* dSrow = dSb + row0 * kPB;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                float dP[4] = {0.f, 0.f, 0.f, 0.f};
                uint32_t bvv[2];
                ldsm_x2(bvv, vs + (j * 8 + (lane & 7)) * kTS + h * kHD + ((lane >> 3) & 1) * 8);
                mma_bf16(dP, ada, bvv);
                S[j][0] *= (dP[0] - rd0); S[j][1] *= (dP[1] - rd0);          // dS = P o (dP - rowdot)
                S[j][2] *= (dP[2] - rd1); S[j][3] *= (dP[3] - rd1);
                *reinterpret_cast<uint32_t*>(dSrow + gq * kPB + j * 8 + tq * 2) = pack_bf16(S[j][0], S[j][1]);
                *reinterpret_cast<uint32_t*>(dSrow + (gq + 8) * kPB + j * 8 + tq * 2) = pack_bf16(S[j][2], S[j][3]);
            }
            *reinterpret_cast<uint32_t*>(dSrow + gq * kPB + 56 + tq * 2) = 0u;
            *reinterpret_cast<uint32_t*>(dSrow + (gq + 8) * kPB + 56 + tq * 2) = 0u;
            // dq = dS k
#pragma unroll
            for (int nd = 0; nd < 2; ++nd)
#pragma unroll
                for (int e = 0; e < 4; ++e) dq[h][nd][e] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t as[4];
                as[0] = pack_bf16(S[2 * kk][0], S[2 * kk][1]);
                as[1] = pack_bf16(S[2 * kk][2], S[2 * kk][3]);
                as[2] = kk < 3 ? pack_bf16(S[kk < 3 ? 2 * kk + 1 : 6][0], S[kk < 3 ? 2 * kk + 1 : 6][1]) : 0u;
                as[3] = kk < 3 ? pack_bf16(S[kk < 3 ? 2 * kk + 1 : 6][2], S[kk < 3 ? 2 * kk + 1 : 6][3]) : 0u;
#pragma unroll
                for (int nd = 0; nd < 2; ++nd) {
                    uint32_t bkk[2];
                    ldsm_x2_t(bkk, ks + (kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kTS + h * kHD + nd * 8);
                    mma_bf16(dq[h][nd], as, bkk);
                }
            }
            __syncthreads();
            // ---- key side of head h: this warp owns keys row0..row0+15: dk = dS^T q, dv = P^T dA ; then the Q^T K gate terms
            float dk[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) {                 // queries 16kq..16kq+15 are the contraction index
                uint32_t a1[4], a2[4];
                const int i = lane >> 3;
                // A[m=key][k=query] = dS[query][key]: transposed load of the [query][key] tile
                ldsm_x4_t(a1, dSb + (kq * 16 + (lane & 7) + 8 * (i >> 1)) * kPB + row0 + 8 * (i & 1));
                ldsm_x4_t(a2, Pb + (kq * 16 + (lane & 7) + 8 * (i >> 1)) * kPB + row0 + 8 * (i & 1));
#pragma unroll
                for (int nd = 0; nd < 2; ++nd) {
                    uint32_t bq[2], bd[2];
                    ldsm_x2_t(bq, qs + (kq * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kTS + h * kHD + nd * 8);
                    ldsm_x2_t(bd, dAb + (kq * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kTS + h * kHD + nd * 8);
                    mma_bf16(dk[nd], a1, bq);
                    mma_bf16(dv[nd], a2, bd);
                }
            }
            const float dz = (misc[4 + h * 4 + 0] + misc[4 + h * 4 + 1] + misc[4 + h * 4 + 2] + misc[4 + h * 4 + 3]) * gate * (1.0f - gate);
            const float u = dz * (1.0f / 256.0f);
            const int am = reinterpret_cast<int*>(misc)[2 + h], as_ = am >> 4, bs_ = am & 15;
            // rows this thread holds in the accumulator layout: row0+gq and row0+gq+8 (queries for dq, keys for dk)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int r = row0 + gq + half * 8;
                float qsum = 0.f, ksum = 0.f;
#pragma unroll
                for (int d = 0; d < 16; ++d) {
                    qsum += __bfloat162float(qs[r * kTS + h * kHD + d]);
                    ksum += __bfloat162float(ks[r * kTS + h * kHD + d]);
                }
                const float q_as = __bfloat162float(qs[r * kTS + h * kHD + as_]), k_bs = __bfloat162float(ks[r * kTS + h * kHD + bs_]);
#pragma unroll
                for (int nd = 0; nd < 2; ++nd)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int d = nd * 8 + tq * 2 + e;
                        dq[h][nd][half * 2 + e] += u * ksum + (d == as_ ? dz * k_bs : 0.f);
                        dk[nd][half * 2 + e] += u * qsum + (d == bs_ ? dz * q_as : 0.f);
                    }
            }
            __syncwarp();                                    // every lane has read its rows of k (now overwritten by dk)
#pragma unroll
            for (int nd = 0; nd < 2; ++nd) {
                const int col = h * kHD + nd * 8 + tq * 2;
                const int r0 = row0 + gq, r1 = r0 + 8;
                *reinterpret_cast<uint32_t*>(Gq + r0 * kTS + col) = pack_bf16(0.25f * dq[h][nd][0], 0.25f * dq[h][nd][1]);
                *reinterpret_cast<uint32_t*>(Gq + r1 * kTS + col) = pack_bf16(0.25f * dq[h][nd][2], 0.25f * dq[h][nd][3]);
                *reinterpret_cast<uint32_t*>(Gk + r0 * kTS + col) = pack_bf16(dk[nd][0], dk[nd][1]);
                *reinterpret_cast<uint32_t*>(Gk + r1 * kTS + col) = pack_bf16(dk[nd][2], dk[nd][3]);
                *reinterpret_cast<uint32_t*>(Gv + r0 * kTS + col) = pack_bf16(dv[nd][0], dv[nd][1]);
                *reinterpret_cast<uint32_t*>(Gv + r1 * kTS + col) = pack_bf16(dv[nd][2], dv[nd][3]);
            }
            __syncthreads();                                 // P / dS of this head are free for the next one; Gq, Gk, Gv complete after h = 1
        }
        // ---- gradients w.r.t. the gated tokens (rows of this warp): dxs = Gq.Wq ; dys = Gk.Wk + Gv.Wv
        {
            float ax[4][4], ay[4][4];
            projT_mma(Gq, Wsm + 0 * kC * kTS, row0, lane, ax, true);
            projT_mma(Gk, Wsm + 1 * kC * kTS, row0, lane, ay, true);
            projT_mma(Gv, Wsm + 2 * kC * kTS, row0, lane, ay, false);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int t = row0 + gq + half * 8;
                const int n = t < kL ? token_pixel(g, wi, wj, t) : -1;
                if (n >= 0) {
                    float* px = dxg + ((size_t)b * g.HW + n) * kC + tq * 2;
                    float* py = dyg + ((size_t)b * g.HW + n) * kC + tq * 2;
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        *reinterpret_cast<float2*>(px + nt * 8) = make_float2(ax[nt][half * 2], ax[nt][half * 2 + 1]);
                        *reinterpret_cast<float2*>(py + nt * 8) = make_float2(ay[nt][half * 2], ay[nt][half * 2 + 1]);
                    }
                }
            }
        }
        // ---- weight gradients: warp m accumulates dW_m[c][i] += sum_t G_m[t][c] X_m[t][i]
        {
            const __nv_bfloat16* G = warp == 0 ? Gq : (warp == 1 ? Gk : (warp == 2 ? Gv : ds));
            const __nv_bfloat16* X = warp == 0 ? xs : (warp == 3 ? Om : ys);
#pragma unroll
            for (int kt = 0; kt < 4; ++kt) {
                uint32_t a[2][4];
                const int i = lane >> 3;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)      // A[m=c][k=t] = G[t][c]: transposed load
                    ldsm_x4_t(a[mt], G + (kt * 16 + (lane & 7) + 8 * (i >> 1)) * kTS + mt * 16 + 8 * (i & 1));
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    uint32_t bx[2];
                    ldsm_x2_t(bx, X + (kt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kTS + nt * 8);
                    mma_bf16(accW[0][nt], a[0], bx);
                    mma_bf16(accW[1][nt], a[1], bx);
                }
            }
            float sb = 0.f;
            for (int t = 0; t < kL; ++t) sb += __bfloat162float(G[t * kTS + lane]);
            accB += sb;
        }
        __syncthreads();
    }
    float* dW = warp == 0 ? gr.q_w : (warp == 1 ? gr.k_w : (warp == 2 ? gr.v_w : gr.o_w));
    float* dB = warp == 0 ? gr.q_b : (warp == 1 ? gr.k_b : (warp == 2 ? gr.v_b : gr.o_b));
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int c0 = mt * 16 + gq, i0 = nt * 8 + tq * 2;
            atomicAdd(dW + c0 * kC + i0, accW[mt][nt][0]);
            atomicAdd(dW + c0 * kC + i0 + 1, accW[mt][nt][1]);
            atomicAdd(dW + (c0 + 8) * kC + i0, accW[mt][nt][2]);
            atomicAdd(dW + (c0 + 8) * kC + i0 + 1, accW[mt][nt][3]);
        }
    atomicAdd(dB + lane, accB);
}

// ------------------------------------------------------------------------------------------
// backward of the saliency gate
// ------------------------------------------------------------------------------------------
// dgmap[b][z][j] = sum_k dgated_flat[k*HW+j] * normed_flat[k*HW+j]
template <typename T>
__global__ void gate_bwd_reduce_kernel(const float* __restrict__ dxg, const float* __restrict__ dyg, const T* __restrict__ x,
                                       const T* __restrict__ y, LnRef lx, LnRef ly, float* __restrict__ dgmap, int HW) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= HW) return;
    const int b = blockIdx.y, z = blockIdx.z;
    const float* d = (z == 0 ? dxg : dyg) + (size_t)b * HW * kC;
    const T* n = (z == 0 ? x : y) + (size_t)b * HW * kC;
    const LnRef ln = z == 0 ? lx : ly;
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < kC; ++k) s += d[(size_t)k * HW + j] * normed_at(n, ln, (size_t)b * HW, (size_t)k * HW + j);
    dgmap[((size_t)b * 2 + z) * HW + j] = s;
}

template <typename T>
__global__ void __launch_bounds__(128) gate_bwd_reduce_vec_kernel(const float* __restrict__ dxg, const float* __restrict__ dyg,
                                                                  const T* __restrict__ x, const T* __restrict__ y, LnRef lx, LnRef ly,
                                                                  float* __restrict__ dgmap, int HW) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;         // 4 lanes per group of 8 positions, rows k split between them
    const int kq = t & 3;
    int j0 = (t >> 2) * 8;
    const bool live = j0 < HW;
    if (!live) j0 = 0;
    const int b = blockIdx.y, z = blockIdx.z;
    const float* d = (z == 0 ? dxg : dyg) + (size_t)b * HW * kC;
    const T* n = (z == 0 ? x : y) + (size_t)b * HW * kC;
    const LnRef ln = z == 0 ? lx : ly;
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
#pragma unroll
    for (int kk = 0; kk < kC / 4; ++kk) {
        const int k = kk * 4 + kq;
        float v[8], g[8];
        load8(d + (size_t)k * HW + j0, g);
        normed8_at(n, ln, (size_t)b * HW, (size_t)k * HW + j0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += g[i] * v[i];
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1)
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
    if (live && kq == 0) store8(dgmap + ((size_t)b * 2 + z) * HW + j0, s);
}

// softmax(2) -> 1x1 conv -> sigmoid backward; writes dpre (grad at the 7x7 conv outputs), accumulates dlvl_w/dlvl_b
__global__ void gate_bwd_map_kernel(const float* __restrict__ dgmap, const float* __restrict__ gmap, const float* __restrict__ smap,
                                    const float* __restrict__ lvl_w, float* __restrict__ dpre,
                                    float* __restrict__ dlvl_w, float* __restrict__ dlvl_b, int HW) {
    __shared__ float red[6][8];
    const int b = blockIdx.y, pix = blockIdx.x * blockDim.x + threadIdx.x;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // dW00 dW01 dW10 dW11 db0 db1
    if (pix < HW) {
        const size_t o0 = ((size_t)b * 2 + 0) * HW + pix, o1 = ((size_t)b * 2 + 1) * HW + pix;
        const float g0 = gmap[o0], g1 = gmap[o1], dg0 = dgmap[o0], dg1 = dgmap[o1];
        const float dot = dg0 * g0 + dg1 * g1;
        const float dl0 = g0 * (dg0 - dot), dl1 = g1 * (dg1 - dot);
        const float s0 = smap[o0], s1 = smap[o1];
        acc[0] = dl0 * s0; acc[1] = dl0 * s1; acc[2] = dl1 * s0; acc[3] = dl1 * s1; acc[4] = dl0; acc[5] = dl1;
        const float ds0 = lvl_w[0] * dl0 + lvl_w[2] * dl1, ds1 = lvl_w[1] * dl0 + lvl_w[3] * dl1;
        dpre[o0] = ds0 * s0 * (1.0f - s0);
        dpre[o1] = ds1 * s1 * (1.0f - s1);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < 6; ++e) { const float r = warp_sum(acc[e]); if (lane == 0) red[e][warp] = r; }
    __syncthreads();
    if (threadIdx.x < 6) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
        atomicAdd(threadIdx.x < 4 ? dlvl_w + threadIdx.x : dlvl_b + (threadIdx.x - 4), s);
    }
}

// 7x7 conv backward on 32x32 tiles: dpooled (data grad) and dw_sa (weight grad, one thread per tap)
__global__ void __launch_bounds__(256) gate_bwd_conv_kernel(const float* __restrict__ dpre, const float* __restrict__ pooled,
                                                            const float* __restrict__ w_sa1, const float* __restrict__ w_sa2,
                                                            float* __restrict__ dpooled, float* __restrict__ dw_sa1,
                                                            float* __restrict__ dw_sa2, int H, int W) {
    constexpr int TS = 32, HS = TS + 6;
    __shared__ float dp[HS][HS + 1];
    __shared__ float pl[2][HS][HS + 1];
    __shared__ float wsm[98];
    const int tiles_x = (W + TS - 1) / TS;
    const int ty0 = (blockIdx.x / tiles_x) * TS, tx0 = (blockIdx.x % tiles_x) * TS;
    const int b = blockIdx.y, z = blockIdx.z, HW = H * W;
    const float* wsrc = z == 0 ? w_sa1 : w_sa2;
    if (threadIdx.x < 98) wsm[threadIdx.x] = wsrc[threadIdx.x];
    for (int idx = threadIdx.x; idx < HS * HS; idx += blockDim.x) {
        const int yy = ty0 + idx / HS - 3, xx = tx0 + idx % HS - 3;
        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
        dp[idx / HS][idx % HS] = in ? dpre[((size_t)b * 2 + z) * HW + yy * W + xx] : 0.f;
        pl[0][idx / HS][idx % HS] = in ? pooled[((size_t)b * 4 + z * 2 + 0) * HW + yy * W + xx] : 0.f;
        pl[1][idx / HS][idx % HS] = in ? pooled[((size_t)b * 4 + z * 2 + 1) * HW + yy * W + xx] : 0.f;
    }
    __syncthreads();
    // data grad: dpooled[ci][y][x] = sum_{dy,dx} dpre[y-dy+3][x-dx+3] * w[ci][dy][dx]
    for (int idx = threadIdx.x; idx < TS * TS; idx += blockDim.x) {
        const int ly = idx / TS, lx = idx % TS, yy = ty0 + ly, xx = tx0 + lx;
        if (yy >= H || xx >= W) continue;
        float a0 = 0.f, a1 = 0.f;
        for (int dy = 0; dy < 7; ++dy)
            for (int dx = 0; dx < 7; ++dx) {
                const float d = dp[ly + 6 - dy][lx + 6 - dx];
                a0 += d * wsm[dy * 7 + dx];
                a1 += d * wsm[49 + dy * 7 + dx];
            }
        dpooled[((size_t)b * 4 + z * 2 + 0) * HW + yy * W + xx] = a0;
        dpooled[((size_t)b * 4 + z * 2 + 1) * HW + yy * W + xx] = a1;
    }
    // weight grad: dw[ci][dy][dx] += sum_{y,x in tile} dpre[y][x] * pooled[ci][y+dy-3][x+dx-3]
    if (threadIdx.x < 98) {
        const int ci = threadIdx.x / 49, dy = (threadIdx.x % 49) / 7, dx = threadIdx.x % 7;
        float a = 0.f;
        for (int ly = 0; ly < TS; ++ly)
            for (int lx = 0; lx < TS; ++lx) a += dp[ly + 3][lx + 3] * pl[ci][ly + dy][lx + dx];
        atomicAdd((z == 0 ? dw_sa1 : dw_sa2) + threadIdx.x, a);
    }
}

// d(normed)[f] = d(gated)[f]*g[f mod HW] + dpooled_avg[j]/C + dpooled_max[j]*[k == argmax[j]],  f = k*HW + j
template <typename T, typename TOUT>
__global__ void gate_bwd_apply_kernel(const float* __restrict__ dxg, const float* __restrict__ dyg, const float* __restrict__ gmap,
                                      const float* __restrict__ dpooled, const uint8_t* __restrict__ amax,
                                      const T* __restrict__ x_add, TOUT* __restrict__ dxn, TOUT* __restrict__ dyn, int HW) {
    const int b = blockIdx.y, z = blockIdx.z;
    const size_t img = (size_t)b * HW * kC;
    const float* src = (z == 0 ? dxg : dyg) + img;
    TOUT* dst = (z == 0 ? dxn : dyn) + img;
    const T* add = (z == 0 && x_add) ? x_add + img : nullptr;
    const float* gm = gmap + ((size_t)b * 2 + z) * HW;
    const float* da = dpooled + ((size_t)b * 4 + z * 2 + 0) * HW;
    const float* dm = dpooled + ((size_t)b * 4 + z * 2 + 1) * HW;
    const uint8_t* am = amax + ((size_t)b * 2 + z) * HW;
    const int64_t total8 = (int64_t)HW * kC / 8;
    for (int64_t i8 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i8 < total8; i8 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f0 = i8 * 8;
        float v[8];
        load8(src + f0, v);
        // f0 < HW*kC: 32-bit division whenever that fits (always, for any image this path sees)
        int kk = (int64_t)HW * kC < 0x7fffffffLL ? (int)((uint32_t)f0 / (uint32_t)HW) : (int)(f0 / HW);
        int j = (int)(f0 - (int64_t)kk * HW);
        if ((HW & 7) == 0) {                 // the 8 positions stay inside one row k of the flat view: vector loads of the maps
            float g8[8], a8[8], m8[8];
            load8(gm + j, g8);
            load8(da + j, a8);
            load8(dm + j, m8);
            const uint2 am8 = __ldg(reinterpret_cast<const uint2*>(am + j));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int a = (int)(((i < 4 ? am8.x : am8.y) >> (8 * (i & 3))) & 0xffu);
                v[i] = v[i] * g8[i] + a8[i] * (1.0f / kC) + (a == kk ? m8[i] : 0.f);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i] = v[i] * gm[j] + da[j] * (1.0f / kC) + ((int)am[j] == kk ? dm[j] : 0.f);
                if (++j == HW) { j = 0; ++kk; }
            }
        }
        if (add) {
            float a[8];
            load8(add + f0, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += a[i];
        }
        store8(dst + f0, v);
    }
}

static inline LnRef ln_ref(const rss_attn_params* p, const float* ln_stats, int64_t rows, int which) {
    if (!p->ln_w) return LnRef{nullptr, nullptr, nullptr, nullptr};
    return LnRef{ln_stats + (2 * which) * rows, ln_stats + (2 * which + 1) * rows, p->ln_w, p->ln_b};
}

template <typename T>
static int attn_fwd_impl(const void* x, const void* y, const rss_attn_params* p, int B, int H, int W, int flags,
                         float* ln_stats, float* pooled, uint8_t* amax, float* smap, float* gmap, void* out, cudaStream_t st) {
    const WinGeom g = make_geom(B, H, W);
    const int64_t rows = (int64_t)B * g.HW;
    const int dt = sizeof(T) == 4 ? RSS_F32 : RSS_BF16;
    int rc;
    if (p->ln_w) {      // norm1 statistics only; the normalised tokens are recomputed where they are consumed
        if ((rc = rss_layernorm_fwd(x, nullptr, ln_stats, ln_stats + rows, p->ln_w, p->ln_b, p->ln_eps, rows, kC, dt, st)) != RSS_OK) return rc;
        if ((rc = rss_layernorm_fwd(y, nullptr, ln_stats + 2 * rows, ln_stats + 3 * rows, p->ln_w, p->ln_b, p->ln_eps, rows, kC, dt, st)) != RSS_OK) return rc;
    }
    const LnRef lx = ln_ref(p, ln_stats, rows, 0), ly = ln_ref(p, ln_stats, rows, 1);
    // (every pointer below is an allocation base + a multiple of HW elements: HW % 8 == 0 keeps the 16/32-byte vectors aligned)
    if (!(flags & RSS_ATTN_NO_GATE)) {      // NO_GATE: the caller filled gmap (Mhca.forward alone: all ones)
        if (g.HW % 8 == 0) {
            dim3 pg((g.HW / 8 * 4 + 127) / 128, B, 2);
            gate_pool_vec_kernel<T><<<pg, 128, 0, st>>>((const T*)x, (const T*)y, lx, ly, pooled, amax, g.HW);
        } else {
            dim3 pg((g.HW + 255) / 256, B, 2);
            gate_pool_kernel<T><<<pg, 256, 0, st>>>((const T*)x, (const T*)y, lx, ly, pooled, amax, g.HW);
        }
        dim3 mg(((W + kGmTW - 1) / kGmTW) * ((H + kGmTH - 1) / kGmTH), B);
        gate_map_kernel<<<mg, kGmTW * kGmTH, 0, st>>>(pooled, p->sa1_w, p->sa2_w, p->lvl_w, p->lvl_b, smap, gmap, H, W);
    }
    const size_t smem = kFwdSmemFloats * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(win_attn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(win_attn_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    int grid = num_sms() * 4;
    if (grid > g.nWin) grid = g.nWin;
    const T* res = (flags & RSS_ATTN_NO_RESIDUAL) ? (const T*)nullptr : (const T*)x;
    if (sizeof(T) == 2 && !(flags & RSS_ATTN_SIMT)) {       // bf16 activations: tensor-core kernel (6 CTAs/SM)
        int gtc = num_sms() * 6;
        if (gtc > g.nWin) gtc = g.nWin;
        win_attn_fwd_tc_kernel<T><<<gtc, kThreads, 0, st>>>((const T*)x, (const T*)y, lx, ly, gmap, res, (T*)out, *p, g);
    } else {
        win_attn_fwd_kernel<T><<<grid, kThreads, smem, st>>>((const T*)x, (const T*)y, lx, ly, gmap, res, (T*)out, *p, g);
    }
    return check_launch();
}

template <typename T>
static int attn_bwd_impl(const void* dout, const void* x, const void* y, const rss_attn_params* p, int B, int H, int W, int flags,
                         const float* ln_stats, const float* pooled, const uint8_t* amax,
                         const float* smap, const float* gmap, void* workspace, void* dx, void* dy, const rss_attn_grads* gr,
                         cudaStream_t st) {
    const WinGeom g = make_geom(B, H, W);
    const int64_t rows = (int64_t)B * g.HW;
    const size_t tok_bytes = (size_t)rows * kC * sizeof(float);   // gradient intermediates stay fp32 (LN backward cancels)
    char* ws = (char*)workspace;
    float* dxg = (float*)ws;  ws += tok_bytes;
    float* dyg = (float*)ws;  ws += tok_bytes;
    float* dgmap = (float*)ws;  ws += (size_t)B * 2 * g.HW * sizeof(float);
    float* dpre = (float*)ws;   ws += (size_t)B * 2 * g.HW * sizeof(float);
    float* dpooled = (float*)ws;
    const bool has_ln = p->ln_w != nullptr, has_res = !(flags & RSS_ATTN_NO_RESIDUAL);
    const LnRef lx = ln_ref(p, ln_stats, rows, 0), ly = ln_ref(p, ln_stats, rows, 1);
    const int dt = sizeof(T) == 4 ? RSS_F32 : RSS_BF16;

    const size_t smem = kBwdSmemFloats * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(win_attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(win_attn_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    int grid = num_sms() * 2;
    if (grid > g.nWin) grid = g.nWin;
    if (sizeof(T) == 2 && !(flags & RSS_ATTN_SIMT)) {
        grid = num_sms() * 3 < g.nWin ? num_sms() * 3 : g.nWin;      // 75.4 KB of shared memory, <= 168 registers: 3 CTAs/SM
        static bool tc_attr = false;
        if (!tc_attr) {
            cudaFuncSetAttribute(win_attn_bwd_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdTcSmem);
            cudaFuncSetAttribute(win_attn_bwd_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdTcSmem);
            tc_attr = true;
        }
        win_attn_bwd_tc_kernel<T><<<grid, kThreads, kBwdTcSmem, st>>>((const T*)x, (const T*)y, lx, ly, gmap, (const T*)dout, dxg, dyg, *p, *gr, g);
    } else {
        win_attn_bwd_kernel<T><<<grid, kThreads, smem, st>>>((const T*)x, (const T*)y, lx, ly, gmap, (const T*)dout, dxg, dyg, *p, *gr, g);
    }
    if (flags & RSS_ATTN_NO_GATE) {         // constant gate (ones): no gradient flows into the gate branch
        cudaError_t e = cudaMemsetAsync(dpooled, 0, (size_t)B * 4 * g.HW * sizeof(float), st);
        if (e != cudaSuccess) { g_last_cuda_error = (int)e; return RSS_ERR_CUDA; }
    } else {
        if (g.HW % 8 == 0) {
            dim3 pg((g.HW / 8 * 4 + 127) / 128, B, 2);
            gate_bwd_reduce_vec_kernel<T><<<pg, 128, 0, st>>>(dxg, dyg, (const T*)x, (const T*)y, lx, ly, dgmap, g.HW);
        } else {
            dim3 pg((g.HW + 255) / 256, B, 2);
            gate_bwd_reduce_kernel<T><<<pg, 256, 0, st>>>(dxg, dyg, (const T*)x, (const T*)y, lx, ly, dgmap, g.HW);
        }
        dim3 mg((g.HW + 255) / 256, B);
        gate_bwd_map_kernel<<<mg, 256, 0, st>>>(dgmap, gmap, smap, p->lvl_w, dpre, gr->lvl_w, gr->lvl_b, g.HW);
        dim3 cg(((H + 31) / 32) * ((W + 31) / 32), B, 2);
        gate_bwd_conv_kernel<<<cg, 256, 0, st>>>(dpre, pooled, p->sa1_w, p->sa2_w, dpooled, gr->sa1_w, gr->sa2_w, H, W);
    }
    int ag = (int)(((int64_t)g.HW * kC / 8 + 255) / 256);
    if (ag > 1024) ag = 1024;
    dim3 apg(ag, B, 2);
    // the apply kernel is elementwise: in place (fp32) when LayerNorm backward follows, else straight into dx/dy
    // with LayerNorm-1 behind it (the block path) and HW % 8 == 0 the apply step rides on the LayerNorm backward's load of the gradient:
    // one launch for gate apply + both LayerNorm backwards (RSS_GATE_LN_FUSED=0: the three separate launches)
    static const bool fuse_ln = !(getenv("RSS_GATE_LN_FUSED") && getenv("RSS_GATE_LN_FUSED")[0] == '0');
    if (has_ln && fuse_ln && (g.HW & 7) == 0 && kC == 32) {
        int rc0 = check_launch();
        if (rc0 != RSS_OK) return rc0;
        return layernorm_bwd_gated(dxg, dyg, x, y, ln_stats, p->ln_w, has_res ? dout : nullptr, dx, dy, gr->ln_w, gr->ln_b, rows, g.HW,
                                   gmap, dpooled, amax, dt, st);
    }
    if (has_ln) gate_bwd_apply_kernel<T, float><<<apg, 256, 0, st>>>(dxg, dyg, gmap, dpooled, amax, (const T*)nullptr, dxg, dyg, g.HW);
    else gate_bwd_apply_kernel<T, T><<<apg, 256, 0, st>>>(dxg, dyg, gmap, dpooled, amax, has_res ? (const T*)dout : (const T*)nullptr,
                                                          (T*)dx, (T*)dy, g.HW);
    int rc = check_launch();
    if (rc != RSS_OK || !has_ln) return rc;
    // LayerNorm1 backward for both streams (shared gamma/beta grads); dx also takes the residual path (MTFM:107)
    if ((rc = layernorm_bwd_f32dy(dxg, x, ln_stats, ln_stats + rows, p->ln_w, has_res ? dout : nullptr, dx, gr->ln_w, gr->ln_b, rows, kC, dt, st)) != RSS_OK) return rc;
    return layernorm_bwd_f32dy(dyg, y, ln_stats + 2 * rows, ln_stats + 3 * rows, p->ln_w, nullptr, dy, gr->ln_w, gr->ln_b, rows, kC, dt, st);
}

}  // namespace rss

using namespace rss;

// SpatialAttention.forward alone (multihead_isa_pool_attention.py:101-115): sigmoid(conv7x7([mean_c(x), max_c(x)])) of an
// NCHW-CONTIGUOUS (B,32,H,W) tensor.  The pooling kernel of the fused path reads token memory through the reference's flat
// (B,C,H,W) view (pool:150-151), which for NCHW-contiguous memory is the plain channel pooling this module defines.
extern "C" int rss_spatial_attention_fwd(const void* x_nchw, const float* conv_w /*(1,2,7,7)*/, float* out /*(B,1,H,W)*/,
                                         float* ws_pooled /*[B][4][HW]*/, uint8_t* ws_amax /*[B][2][HW]*/, float* ws_maps /*[2][B][2][HW]*/,
                                         int B, int H, int W, int dtype, cudaStream_t st) {
    if (B <= 0 || H <= 0 || W <= 0 || !x_nchw || !conv_w || !out) return RSS_ERR_SHAPE;
    const int HW = H * W;
    const LnRef none{nullptr, nullptr, nullptr, nullptr};
    float* smap = ws_maps;
    float* gmap = ws_maps + (size_t)B * 2 * HW;
    dim3 pg((HW + 255) / 256, B, 2);
    if (dtype == RSS_F32) gate_pool_kernel<float><<<pg, 256, 0, st>>>((const float*)x_nchw, (const float*)x_nchw, none, none, ws_pooled, ws_amax, HW);
    else if (dtype == RSS_BF16) gate_pool_kernel<__nv_bfloat16><<<pg, 256, 0, st>>>((const __nv_bfloat16*)x_nchw, (const __nv_bfloat16*)x_nchw, none, none, ws_pooled, ws_amax, HW);
    else return RSS_ERR_DTYPE;
    dim3 mg(((W + kGmTW - 1) / kGmTW) * ((H + kGmTH - 1) / kGmTH), B);
    // level weights are irrelevant for smap; the conv weight doubles as a valid (2,2)+(2) parameter block of finite numbers
    gate_map_kernel<<<mg, kGmTW * kGmTH, 0, st>>>(ws_pooled, conv_w, conv_w, conv_w, conv_w, smap, gmap, H, W);
    cudaError_t e = cudaMemcpy2DAsync(out, (size_t)HW * sizeof(float), smap, (size_t)2 * HW * sizeof(float), (size_t)HW * sizeof(float), B,
                                      cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return RSS_ERR_CUDA; }
    return check_launch();
}

extern "C" size_t rss_attn_bwd_workspace_bytes(int B, int H, int W, int dtype) {
    (void)dtype;
    const size_t HW = (size_t)H * W;
    return 2 * (size_t)B * HW * kC * sizeof(float) + (size_t)B * HW * (2 + 2 + 4) * sizeof(float) + 256;
}

extern "C" int rss_attn_fwd(const void* x, const void* y, const rss_attn_params* p, int B, int H, int W, int dtype, int flags,
                            float* ln_stats, float* pooled, uint8_t* amax, float* smap, float* gmap,
                            void* out, cudaStream_t stream) {
    if (B <= 0 || H <= 0 || W <= 0 || !p) return RSS_ERR_SHAPE;
    if (p->C != kC || p->num_heads != 2 || p->window != kWS) return RSS_ERR_SHAPE;
    if (((int64_t)H * W * kC) % 8 || (int64_t)H * W * kC >= 0x7fffffffLL) return RSS_ERR_SHAPE;     // 32-bit per-image index math
    RSS_DISPATCH_DTYPE(dtype, return attn_fwd_impl<T>(x, y, p, B, H, W, flags, ln_stats, pooled, amax, smap, gmap, out, stream));
}

extern "C" int rss_attn_bwd(const void* dout, const void* x, const void* y, const rss_attn_params* p, int B, int H, int W, int dtype, int flags,
                            const float* ln_stats, const float* pooled, const uint8_t* amax,
                            const float* smap, const float* gmap, void* workspace, size_t workspace_bytes,
                            void* dx, void* dy, const rss_attn_grads* grads, cudaStream_t stream) {
    if (B <= 0 || H <= 0 || W <= 0 || !p || !grads) return RSS_ERR_SHAPE;
    if (p->C != kC || p->num_heads != 2 || p->window != kWS) return RSS_ERR_SHAPE;
    if ((int64_t)H * W * kC >= 0x7fffffffLL) return RSS_ERR_SHAPE;
    if (workspace_bytes < rss_attn_bwd_workspace_bytes(B, H, W, dtype)) return RSS_ERR_WORKSPACE;
    RSS_DISPATCH_DTYPE(dtype, return attn_bwd_impl<T>(dout, x, y, p, B, H, W, flags, ln_stats, pooled, amax, smap, gmap,
                                                        workspace, dx, dy, grads, stream));
}
