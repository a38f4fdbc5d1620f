// HRNet stem, first convolution: Conv2d(3, 64, 3, stride 2, padding 1, bias=False) straight from the model input
// (_hrnet_rssformer.py:467-470; the reference feeds it the (B,3,H,W) fp32 image batch).
//
// Why its own kernel: with 3 input channels the library path is five launches -- fp32 -> bf16 cast, NCHW -> NHWC permute, a
// 3 -> 8 channel padding kernel, a legacy (sm80 mma) implicit GEMM, and in the backward pass another padding + split-K + reduction
// chain -- 224 us forward + 160 us weight gradient on the B=16 step, all of it on the one-stream head / tail of the step
// (profiles/timeline_r2_final_526_summary.txt), for a layer whose compulsory traffic is a 50 MB read and a 134 MB write.
//
//   forward   y[b,oy,ox,n] = sum_{ky,kx,c} x[b,c,2oy+ky-1,2ox+kx-1] * w[n,c,ky,kx]          (bf16 NHWC out, fp32 accumulation)
//             + optionally the BatchNorm raw sums  sum (y-K), sum (y-K)^2  of the bf16-rounded outputs (K = the running mean):
//             the statistics pass over the 134 MB output disappears (ops.RAW_SUMS protocol of bn.cu)
//   wgrad     dw[n,c,ky,kx] += sum_{b,oy,ox} dy[b,oy,ox,n] * x[b,c,2oy+ky-1,2ox+kx-1]       (fp32 atomics into the caller's buffer)
//   (no data gradient: the image does not require one)
//
// Both are HBM-bound GEMMs with K = 27 (forward) / N = 27 (weight gradient).  A CTA owns tiles of 128 output pixels of one output
// row.  Per tile the 128 x 27 patch matrix is built ONCE in shared memory (im2col of the planar fp32/bf16 input: the cast and the
// layout change are fused into this gather; each input element is fetched through L1/L2, HBM sees the image once) and fed to
// mma.sync.m16n8k16 (bf16 x bf16 -> fp32) through ldmatrix; K is padded to 32 with zeros.  tcgen05 would not help: the tile's
// 128 x 64 x 32 MACs are ~1 % of the time it takes to write its 16 KB of output.
// Algorithmic bytes per launch (B=16, 512x512): forward 50.3 MB (fp32 image) + 134.2 MB (output) = 184.5 MB;
// weight gradient 134.2 MB (dy) + 50.3 MB = 184.5 MB.
#include "common.cuh"

namespace rss {

constexpr int kStemN = 64;            // output channels
constexpr int kStemK = 27;            // 3 x 3 x 3 taps*channels, k = ky*9 + kx*3 + c
constexpr int kStemKP = 32;           // padded to two k16 steps
constexpr int kStemPix = 128;         // output pixels per tile
constexpr int kStemThreads = 256;     // 8 warps x 16 pixels
constexpr int kStemPP = 40;           // bf16 pitch of a patch row (80 B: conflict-free ldmatrix, 16-byte aligned)
constexpr int kStemOP = 72;           // bf16 pitch of a staged output / dy row (144 B)

struct StemGeom { int B, H, W, Ho, Wo, tiles_x, n_tiles; };

__device__ __forceinline__ float stem_in(const float* p) { return __ldg(p); }
__device__ __forceinline__ float stem_in(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ uint32_t stem_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void stem_ldsm(uint32_t r[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void stem_ldsm_t(uint32_t r[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void stem_mma(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t stem_pack(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// (ky, kx, c) of patch column k and the offset of that tap in a (n, c, ky, kx) weight row
__device__ __forceinline__ void stem_tap(int k, int& ky, int& kx, int& c) {
    ky = k / 9; const int r = k - ky * 9; kx = r / 3; c = r - kx * 3;
}

// P[p][k] (bf16, pitch kStemPP) = x[b, c, 2oy+ky-1, 2(ox0+p)+kx-1], zero outside the image / beyond the row / for k >= 27.
// One k per warp-iteration (128 % 32 == 0): the tap decode is warp-uniform and the 32 lanes read a stride-2 run of one image row.
template <typename TIn>
__device__ __forceinline__ void stem_im2col(const TIn* __restrict__ x, const StemGeom& g, int b, int oy, int ox0,
                                            __nv_bfloat16* __restrict__ P) {
#pragma unroll 4
    for (int idx = threadIdx.x; idx < kStemKP * kStemPix; idx += kStemThreads) {
        const int k = idx / kStemPix, p = idx % kStemPix;
        float v = 0.f;
        if (k < kStemK) {
            int ky, kx, c;
            stem_tap(k, ky, kx, c);
            const int iy = 2 * oy + ky - 1, ix = 2 * (ox0 + p) + kx - 1;
            if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W && ox0 + p < g.Wo)
                v = stem_in(x + (((int64_t)b * 3 + c) * g.H + iy) * g.W + ix);
        }
        P[p * kStemPP + k] = __float2bfloat16_rn(v);
    }
}

// The same matrix when W is even (every image row starts 8-byte aligned): two threads per output pixel, each taking every other
// one of the 9 (ky, c) image rows.  A thread loads the PAIR x[.., 2ox], x[.., 2ox+1] -- taps kx = 1 and kx = 2 of its own pixel and,
// the second one, tap kx = 0 of the pixel to its right -- so every input element of the tile is fetched exactly once, coalesced,
// in ~10 instructions per pair (the generic gather above decodes (ky,kx,c) and bounds per element: ~50 instructions per element
// made BOTH kernels issue-bound at 147 us, 5x their HBM time).  Fetch and store are separate so that the loads of tile i+1 are in
// flight while tile i is multiplied (register double buffer).  Columns 27..31 of P are zeroed once per kernel.
__device__ __forceinline__ void stem_ldpair(const float* p, float& a, float& b) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p)); a = v.x; b = v.y;
}
__device__ __forceinline__ void stem_ldpair(const __nv_bfloat16* p, float& a, float& b) {
    const uint32_t v = __ldg(reinterpret_cast<const unsigned int*>(p));
    a = __uint_as_float(v << 16); b = __uint_as_float(v & 0xffff0000u);
}
struct StemTile { int b, oy, ox0; };
__device__ __forceinline__ StemTile stem_tile(const StemGeom& g, int tile) {
    StemTile t;
    const int tx = tile % g.tiles_x, row = tile / g.tiles_x;
    t.oy = row % g.Ho; t.b = row / g.Ho; t.ox0 = tx * kStemPix;
    return t;
}
struct StemRegs { float v0[5], v1[5], edge; };
template <typename TIn>
__device__ __forceinline__ void stem_fetch(const TIn* __restrict__ x, const StemGeom& g, const StemTile& t, StemRegs& R) {
    const int p = threadIdx.x & (kStemPix - 1), h = threadIdx.x >> 7;
    const int ix = 2 * (t.ox0 + p);
    const TIn* base = x + ((int64_t)t.b * 3 * g.H + 2 * t.oy - 1) * g.W + ix;      // (c = 0, ky = 0); may point one row above the image
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int r = h + 2 * j, ky = r / 3, c = r - ky * 3, iy = 2 * t.oy + ky - 1;      // r = ky*3 + c
        R.v0[j] = 0.f; R.v1[j] = 0.f;
        if (r < 9 && iy >= 0 && iy < g.H && ix + 1 < g.W)                          // W even: ix < W implies ix + 1 < W
            stem_ldpair(base + ((int64_t)c * g.H + ky) * g.W, R.v0[j], R.v1[j]);
    }
    R.edge = 0.f;
    if (threadIdx.x < 9) {                                // tap kx = 0 of the tile's first pixel
        const int ky = threadIdx.x / 3, c = threadIdx.x - ky * 3, iy = 2 * t.oy + ky - 1, ixl = 2 * t.ox0 - 1;
        if (iy >= 0 && iy < g.H && ixl >= 0) R.edge = stem_in(x + (((int64_t)t.b * 3 + c) * g.H + iy) * g.W + ixl);
    }
}
__device__ __forceinline__ void stem_store(const StemRegs& R, __nv_bfloat16* __restrict__ P) {
    const int p = threadIdx.x & (kStemPix - 1), h = threadIdx.x >> 7;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int r = h + 2 * j, ky = r / 3, c = r - ky * 3;
        if (r < 9) {
            const __nv_bfloat16 b0 = __float2bfloat16_rn(R.v0[j]), b1 = __float2bfloat16_rn(R.v1[j]);
            __nv_bfloat16* dst = P + p * kStemPP + ky * 9 + c;
            dst[3] = b0;                                  // kx = 1
            dst[6] = b1;                                  // kx = 2
            if (p + 1 < kStemPix) dst[kStemPP] = b1;      // kx = 0 of the pixel to the right
        }
    }
    if (threadIdx.x < 9) P[(threadIdx.x / 3) * 9 + threadIdx.x % 3] = __float2bfloat16_rn(R.edge);
}
__device__ __forceinline__ void stem_zero_pad_columns(__nv_bfloat16* P) {
    for (int i = threadIdx.x; i < kStemPix * (kStemKP - kStemK); i += kStemThreads)
        P[(i / (kStemKP - kStemK)) * kStemPP + kStemK + i % (kStemKP - kStemK)] = __float2bfloat16_rn(0.f);
}

template <typename TIn>
__global__ void __launch_bounds__(kStemThreads, 3)
stem_conv_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ w /*(64,3,3,3) fp32 master*/, __nv_bfloat16* __restrict__ y,
                     StemGeom g, float* __restrict__ stat_accum /*NULL or [2][64]*/, const float* __restrict__ stat_shift /*NULL or [64]*/) {
    __shared__ __align__(16) __nv_bfloat16 P[kStemPix * kStemPP];     // 10 KB patch matrix
    __shared__ __align__(16) __nv_bfloat16 Wt[kStemN * kStemPP];      // 5 KB weights [n][k]
    __shared__ __align__(16) __nv_bfloat16 O[kStemPix * kStemOP];     // 18 KB output staging (private 16-row slab per warp)
    __shared__ float red[2 * kStemN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
    for (int i = tid; i < kStemN * kStemKP; i += kStemThreads) {
        const int n = i / kStemKP, k = i % kStemKP;
        float v = 0.f;
        if (k < kStemK) { int ky, kx, c; stem_tap(k, ky, kx, c); v = w[n * kStemK + c * 9 + ky * 3 + kx]; }
        Wt[n * kStemPP + k] = __float2bfloat16_rn(v);
    }
    if (tid < 2 * kStemN) red[tid] = 0.f;
    stem_zero_pad_columns(P);
    const bool even = (g.W & 1) == 0 && ((uintptr_t)x & 7) == 0;
    // BatchNorm sums of this lane's 8 channels (cp*8 .. cp*8+7, cp = lane & 7: the 16-byte chunk it copies out of every row),
    // UNshifted here (sum y, sum y^2 over <= ~30 tiles) and moved to the shift K when the CTA publishes them
    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    float cnt = 0.f;                                       // pixels this lane summed
    const int p0 = warp * 16;
    StemRegs R;
    if (even && (int)blockIdx.x < g.n_tiles) stem_fetch(x, g, stem_tile(g, blockIdx.x), R);
    for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const StemTile t = stem_tile(g, tile);
        __syncthreads();                                  // the previous tile's readers of P are done (first pass: Wt / red / padding visible)
        if (even) {
            stem_store(R, P);
            const int next = tile + gridDim.x;
            if (next < g.n_tiles) stem_fetch(x, g, stem_tile(g, next), R);
        } else {
            stem_im2col(x, g, t.b, t.oy, t.ox0, P);
        }
        __syncthreads();
        uint32_t a[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)                    // matrices: (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15)
            stem_ldsm(a[ks], stem_smem(P + (p0 + (lane & 7) + ((lane >> 3) & 1) * 8) * kStemPP + ks * 16 + (lane >> 4) * 8));
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t bw[4];                               // matrices: rows n0..n0+7, k 0-7 / 8-15 / 16-23 / 24-31  ->  b0,b1 of k-step 0, b0,b1 of k-step 1
            stem_ldsm(bw, stem_smem(Wt + (nt * 8 + (lane & 7)) * kStemPP + (lane >> 3) * 8));
            stem_mma(acc, a[0], bw[0], bw[1]);
            stem_mma(acc, a[1], bw[2], bw[3]);
            *reinterpret_cast<uint32_t*>(O + (p0 + gq) * kStemOP + nt * 8 + 2 * tq) = stem_pack(acc[0], acc[1]);
            *reinterpret_cast<uint32_t*>(O + (p0 + gq + 8) * kStemOP + nt * 8 + 2 * tq) = stem_pack(acc[2], acc[3]);
        }
        __syncwarp();
        // 16 rows x 128 B of this warp's slab -> global, 16 bytes per lane and instruction (4 rows of 128 B per instruction); the
        // statistics are taken from the same registers, i.e. from the bf16-rounded values the consumer will read
        __nv_bfloat16* yrow = y + (((int64_t)t.b * g.Ho + t.oy) * g.Wo + t.ox0 + p0) * kStemN;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = (lane >> 3) + 4 * i, cp = lane & 7;
            if (t.ox0 + p0 + r < g.Wo) {
                const uint4 v = *reinterpret_cast<const uint4*>(O + (p0 + r) * kStemOP + cp * 8);
                *reinterpret_cast<uint4*>(yrow + r * kStemN + cp * 8) = v;
                if (stat_accum) {
                    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float lo = __uint_as_float(wv[q] << 16), hi = __uint_as_float(wv[q] & 0xffff0000u);
                        s1[2 * q] += lo; s2[2 * q] = fmaf(lo, lo, s2[2 * q]);
                        s1[2 * q + 1] += hi; s2[2 * q + 1] = fmaf(hi, hi, s2[2 * q + 1]);
                    }
                    cnt += 1.f;
                }
            }
        }
        __syncwarp();
    }
    if (!stat_accum) return;
    // sum (y-K) = S1 - nK ;  sum (y-K)^2 = S2 - 2K S1 + nK^2   (per lane, before the lanes are combined: n is this lane's count)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float K = stat_shift ? stat_shift[(lane & 7) * 8 + i] : 0.f;
        const float a1 = s1[i] - cnt * K;
        const float a2 = s2[i] - 2.f * K * s1[i] + cnt * K * K;
        s1[i] = a1; s2[i] = a2;
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) {
            s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
            s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], o);
        }
    }
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            atomicAdd(&red[lane * 8 + i], s1[i]);
            atomicAdd(&red[kStemN + lane * 8 + i], s2[i]);
        }
    }
    __syncthreads();
    if (tid < 2 * kStemN) atomicAdd(stat_accum + tid, red[tid]);
}

template <typename TIn>
__global__ void __launch_bounds__(kStemThreads, 2)
stem_conv_wgrad_kernel(const TIn* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw /*(64,3,3,3) fp32, +=*/,
                       StemGeom g) {
    __shared__ __align__(16) __nv_bfloat16 P[kStemPix * kStemPP];     // patch matrix [pix][k]
    __shared__ __align__(16) __nv_bfloat16 D[kStemPix * kStemOP];     // dy tile [pix][n]
    __shared__ float red[kStemN * kStemKP];                           // 8 KB: cross-warp sum of dw[n][k]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
    for (int i = tid; i < kStemN * kStemKP; i += kStemThreads) red[i] = 0.f;
    stem_zero_pad_columns(P);
    const bool even = (g.W & 1) == 0 && ((uintptr_t)x & 7) == 0;
    float acc[4][4][4];                                   // [16-channel block of n][8-column block of k][c fragment]
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
    const int p0 = warp * 16;                             // this warp's 16 pixels = one k16 step of the (n x k) += dy^T P product
    StemRegs R;
    uint4 dv[4];
    auto fetch_dy = [&](const StemTile& t) {              // 128 pixels x 128 B, 16 B per thread and iteration
        const __nv_bfloat16* drow = dy + (((int64_t)t.b * g.Ho + t.oy) * g.Wo + t.ox0) * kStemN;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int chunk = tid + kStemThreads * i, r = chunk >> 3, cp = chunk & 7;
            dv[i] = make_uint4(0u, 0u, 0u, 0u);
            if (t.ox0 + r < g.Wo) dv[i] = __ldg(reinterpret_cast<const uint4*>(drow + r * kStemN + cp * 8));
        }
    };
    if ((int)blockIdx.x < g.n_tiles) {
        const StemTile t = stem_tile(g, blockIdx.x);
        fetch_dy(t);
        if (even) stem_fetch(x, g, t, R);
    }
    for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const StemTile t = stem_tile(g, tile);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int chunk = tid + kStemThreads * i;
            *reinterpret_cast<uint4*>(D + (chunk >> 3) * kStemOP + (chunk & 7) * 8) = dv[i];
        }
        if (even) stem_store(R, P);
        else stem_im2col(x, g, t.b, t.oy, t.ox0, P);
        const int next = tile + gridDim.x;
        if (next < g.n_tiles) {                           // next tile's loads in flight while this one is multiplied
            const StemTile tn = stem_tile(g, next);
            fetch_dy(tn);
            if (even) stem_fetch(x, g, tn, R);
        }
        __syncthreads();
        uint32_t bf[2][4];                                // per 16 patch columns: (pix 0-7, k0..7), (pix 8-15, k0..7), (pix 0-7, k0+8..), (pix 8-15, k0+8..)
#pragma unroll
        for (int np = 0; np < 2; ++np)
            stem_ldsm_t(bf[np], stem_smem(P + (p0 + (lane & 7) + ((lane >> 3) & 1) * 8) * kStemPP + np * 16 + (lane >> 4) * 8));
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            uint32_t a[4];                                // (pix 0-7, n0..7), (pix 0-7, n0+8..15), (pix 8-15, n0..7), (pix 8-15, n0+8..15)
            stem_ldsm_t(a, stem_smem(D + (p0 + (lane & 7) + (lane >> 4) * 8) * kStemOP + mt * 16 + ((lane >> 3) & 1) * 8));
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) stem_mma(acc[mt][nt], a, bf[nt >> 1][(nt & 1) * 2], bf[nt >> 1][(nt & 1) * 2 + 1]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int n = mt * 16 + gq, k = nt * 8 + 2 * tq;
            atomicAdd(&red[n * kStemKP + k], acc[mt][nt][0]);
            atomicAdd(&red[n * kStemKP + k + 1], acc[mt][nt][1]);
            atomicAdd(&red[(n + 8) * kStemKP + k], acc[mt][nt][2]);
            atomicAdd(&red[(n + 8) * kStemKP + k + 1], acc[mt][nt][3]);
        }
    __syncthreads();
    for (int e = tid; e < kStemN * kStemK; e += kStemThreads) {
        const int n = e / kStemK, k = e % kStemK;
        int ky, kx, c;
        stem_tap(k, ky, kx, c);
        atomicAdd(dw + n * kStemK + c * 9 + ky * 3 + kx, red[n * kStemKP + k]);
    }
}

static inline bool stem_geom(int B, int H, int W, StemGeom& g) {
    if (B <= 0 || H <= 0 || W <= 0) return false;
    g.B = B; g.H = H; g.W = W;
    g.Ho = (H - 1) / 2 + 1; g.Wo = (W - 1) / 2 + 1;       // floor((H + 2 - 3) / 2) + 1
    g.tiles_x = (g.Wo + kStemPix - 1) / kStemPix;
    const int64_t n = (int64_t)B * g.Ho * g.tiles_x;
    if (n > 0x7fffffff) return false;
    g.n_tiles = (int)n;
    return true;
}

}  // namespace rss

using namespace rss;

// x: (B,3,H,W) planar, in_dtype RSS_F32 or RSS_BF16; w: (64,3,3,3) fp32; y: (B,Ho,Wo,64) bf16 NHWC, Ho = (H-1)/2+1, Wo = (W-1)/2+1.
// stat_accum (optional, [2][64] fp32, must hold zeros or earlier partial sums): += sum (y-K), sum (y-K)^2 per channel over the
// bf16-rounded outputs, K = stat_shift[c] (NULL: 0) -- the raw sums rss_bn_act_fwd_raw finalises.
extern "C" int rss_stem_conv_fwd(const void* x, const float* w, void* y, int B, int H, int W, int in_dtype,
                                 float* stat_accum, const float* stat_shift, cudaStream_t st) {
    StemGeom g;
    if (!stem_geom(B, H, W, g)) return RSS_ERR_SHAPE;
    if (((uintptr_t)y & 15) || ((uintptr_t)w & 3)) return RSS_ERR_SHAPE;
    int grid = g.n_tiles < num_sms() * 3 ? g.n_tiles : num_sms() * 3;
    if (in_dtype == RSS_F32)
        stem_conv_fwd_kernel<float><<<grid, kStemThreads, 0, st>>>((const float*)x, w, (__nv_bfloat16*)y, g, stat_accum, stat_shift);
    else if (in_dtype == RSS_BF16)
        stem_conv_fwd_kernel<__nv_bfloat16><<<grid, kStemThreads, 0, st>>>((const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, g, stat_accum, stat_shift);
    else
        return RSS_ERR_DTYPE;
    return check_launch();
}

// dw_acc (64,3,3,3) fp32 += weight gradient; x as above, dy (B,Ho,Wo,64) bf16 NHWC.
extern "C" int rss_stem_conv_wgrad(const void* x, const void* dy, float* dw_acc, int B, int H, int W, int in_dtype, cudaStream_t st) {
    StemGeom g;
    if (!stem_geom(B, H, W, g)) return RSS_ERR_SHAPE;
    if ((uintptr_t)dy & 15) return RSS_ERR_SHAPE;
    int grid = g.n_tiles < num_sms() * 2 ? g.n_tiles : num_sms() * 2;
    if (in_dtype == RSS_F32)
        stem_conv_wgrad_kernel<float><<<grid, kStemThreads, 0, st>>>((const float*)x, (const __nv_bfloat16*)dy, dw_acc, g);
    else if (in_dtype == RSS_BF16)
        stem_conv_wgrad_kernel<__nv_bfloat16><<<grid, kStemThreads, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw_acc, g);
    else
        return RSS_ERR_DTYPE;
    return check_launch();
}
