// Weight gradient of the HRNet-family convolutions (NHWC bf16 activations, fp32 accumulation straight into the
// caller's fp32 gradient buffer): BasicBlock/Bottleneck 3x3 and 1x1 convs, the stride-2 3x3 convs of the
// transition/fuse layers, the 1x1 fuse convs (_hrnet_rssformer.py:209-287,361-405,512-546) and the FFN's fc1/fc2
// (ffn_block.py:219,232).
//
//   dW[co][ci][ky][kx] += sum_{b,oy,ox} dY[b,oy,ox,co] * X[b, oy*s + ky*d - p, ox*s + kx*d - p, ci]
//
// GEMM view per tap: M = Cout, N = Cin, K = B*Ho*Wo output pixels (split-K over CTAs).  These layers have tiny M,N
// (32..256) and huge K (up to 262144), i.e. they are HBM/L2-bound: each CTA owns one (32 cout x 32 cin) block of ALL
// k*k taps, walks over 2-D spatial tiles of 256 (stride 1) / 128 (stride 2) output pixels, stages the dY tile and the
// X tile WITH HALO once per tile in shared memory (cp.async, double buffered, zero fill outside the image = padding),
// and feeds the 9 taps from the same staged X tile at shifted ldmatrix row addresses, so X and dY are read from
// L2/HBM once per (cout-block, cin-block) pair instead of once per tap.  Both operands are "transposed" for the MMA
// (pixels are the K dimension but channels are contiguous in memory): ldmatrix.trans builds the fragments.
// Math: mma.sync.m16n8k16 bf16 x bf16 -> fp32 (the K=pixels operands are MN-major; this kernel is bandwidth-bound,
// see DESIGN.md).  Per-CTA accumulators are flushed once at the end through shared memory as 16-byte vector
// reductions (red.global.add.v4.f32) into the PyTorch (Cout,Cin,k,k) layout.
#include "common.cuh"

namespace rss {

constexpr int kWgThreads = 128;      // 4 warps: 2 (cout halves) x 2 (cin halves) of the 32x32 block
constexpr int kWgBlk = 32;           // channels per block (both cout and cin)
constexpr int kWgPitch = 40;         // bf16 elements per staged pixel row (80 B: conflict-free ldmatrix)

struct WgGeom {
    int B, Hi, Wi, Cin, Ho, Wo, Cout;
    int pad, dil;
    int TH, TW;                      // output tile (TW in {16, 32}); TH*TW = 256 (stride 1) or 128 (stride 2)
    int IH, IW;                      // staged input tile incl. halo
    int tiles_x, tiles_y, n_tiles;
    int cin_blocks;
};

__device__ __forceinline__ void wg_mma(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void wg_ldsm_t(uint32_t r[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void wg_cp16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;     // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void wg_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wg_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void wg_red4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// KS in {1,3}; S in {1,2}
template <int KS, int S>
__global__ void __launch_bounds__(kWgThreads)
conv_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw, WgGeom g) {
    constexpr int TAPS = KS * KS;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const int co0 = (blockIdx.y / g.cin_blocks) * kWgBlk, ci0 = (blockIdx.y % g.cin_blocks) * kWgBlk;
    const int in_pix = g.IH * g.IW, out_pix = g.TH * g.TW;
    const uint32_t x_bytes = (uint32_t)in_pix * kWgPitch * 2, stage_bytes = x_bytes + (uint32_t)out_pix * kWgPitch * 2;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);

    float acc[TAPS][2][4];
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[t][j][i] = 0.f;

    auto stage_tile = [&](int tile, int buf) {
        const int tx = tile % g.tiles_x, ty = (tile / g.tiles_x) % g.tiles_y, b = tile / (g.tiles_x * g.tiles_y);
        const int oy0 = ty * g.TH, ox0 = tx * g.TW;
        const int iy0 = oy0 * S - g.pad, ix0 = ox0 * S - g.pad;
        const uint32_t xs = sbase + buf * stage_bytes, ds = xs + x_bytes;
        const int q = tid & 3;                                         // 4 x 16 B per staged pixel; 32 pixels per pass
        {
            int ry = 0, cx = tid >> 2;
            while (cx >= g.IW) { cx -= g.IW; ++ry; }
            for (int p = tid >> 2; p < in_pix; p += kWgThreads / 4) {
                const int iy = iy0 + ry, ix = ix0 + cx;
                const bool ok = iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi;
                const __nv_bfloat16* src = ok ? x + (((size_t)b * g.Hi + iy) * g.Wi + ix) * g.Cin + ci0 + q * 8 : x;
                wg_cp16(xs + (uint32_t)(p * kWgPitch + q * 8) * 2, src, ok);
                cx += kWgThreads / 4;
                while (cx >= g.IW) { cx -= g.IW; ++ry; }
            }
        }
        {
            int ry = 0, cx = tid >> 2;
            while (cx >= g.TW) { cx -= g.TW; ++ry; }
            for (int p = tid >> 2; p < out_pix; p += kWgThreads / 4) {
                const int oy = oy0 + ry, ox = ox0 + cx;
                const bool ok = oy < g.Ho && ox < g.Wo;
                const __nv_bfloat16* src = ok ? dy + (((size_t)b * g.Ho + oy) * g.Wo + ox) * g.Cout + co0 + q * 8 : dy;
                wg_cp16(ds + (uint32_t)(p * kWgPitch + q * 8) * 2, src, ok);
                cx += kWgThreads / 4;
                while (cx >= g.TW) { cx -= g.TW; ++ry; }
            }
        }
    };

    // ldmatrix lane roles (see the fragment layouts of mma.m16n8k16):
    //   A = dY^T (m = cout, k = pixel), stored [pixel][cout]: matrices (m0-7,k0-7) (m8-15,k0-7) (m0-7,k8-15) (m8-15,k8-15)
    //   B = X     (k = pixel, n = cin), stored [pixel][cin] : matrices (k0-7,n0-7) (k8-15,n0-7) (k0-7,n8-15) (k8-15,n8-15)
    const int a_k = (lane & 7) + ((lane >> 4) << 3), a_m = wm * 16 + ((lane >> 3) & 1) * 8;
    const int b_k = (lane & 7) + (((lane >> 3) & 1) << 3), b_n = wn * 16 + (lane >> 4) * 8;
    const uint32_t a_lane = (uint32_t)(a_k * kWgPitch + a_m) * 2;
    const uint32_t b_lane = (uint32_t)(b_k * S * kWgPitch + b_n) * 2;
    const int chunks_per_row = g.TW >> 4, n_chunks = out_pix >> 4;

    int buf = 0;
    int tile = blockIdx.x;
    if (tile < g.n_tiles) stage_tile(tile, 0);
    wg_commit();
    for (; tile < g.n_tiles; tile += gridDim.x) {
        const int next = tile + gridDim.x;
        if (next < g.n_tiles) stage_tile(next, buf ^ 1);
        wg_commit();
        wg_wait<1>();
        __syncthreads();
        const uint32_t xs = sbase + buf * stage_bytes, ds = xs + x_bytes;
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
            const int r = c / chunks_per_row, h = c % chunks_per_row;
            uint32_t a[4];
            wg_ldsm_t(a, ds + (uint32_t)(c * 16 * kWgPitch) * 2 + a_lane);
            const uint32_t xrow = xs + (uint32_t)((r * S * g.IW + h * 16 * S) * kWgPitch) * 2 + b_lane;
#pragma unroll
            for (int ky = 0; ky < KS; ++ky)
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    uint32_t bfr[4];
                    wg_ldsm_t(bfr, xrow + (uint32_t)((ky * g.dil * g.IW + kx * g.dil) * kWgPitch) * 2);
                    wg_mma(acc[ky * KS + kx][0], a, bfr[0], bfr[1]);
                    wg_mma(acc[ky * KS + kx][1], a, bfr[2], bfr[3]);
                }
        }
        __syncthreads();
        buf ^= 1;
    }
    wg_wait<0>();
    __syncthreads();

    // flush: registers -> smem [32 co][32 ci][TAPS] (the PyTorch order of this block) -> 16-byte vector reductions
    float* st = reinterpret_cast<float*>(smem_raw);
    const int gid = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = wm * 16 + gid + (i >> 1) * 8, n = wn * 16 + j * 8 + tig * 2 + (i & 1);
                st[(m * kWgBlk + n) * TAPS + t] = acc[t][j][i];
            }
    __syncthreads();
    constexpr int ROW = kWgBlk * TAPS;                 // contiguous floats per cout row of this block (32*9 or 32)
    for (int i = tid * 4; i < kWgBlk * ROW; i += kWgThreads * 4) {
        const int m = i / ROW, off = i % ROW;
        float* dst = dw + ((size_t)(co0 + m) * g.Cin + ci0) * TAPS + off;
        wg_red4(dst, st[i], st[i + 1], st[i + 2], st[i + 3]);
    }
}

}  // namespace rss

using namespace rss;

// geometry this kernel accepts: bf16 NHWC, k in {1,3}, stride in {1,2}, dilation 1, pad == k/2, Cin % 32 == Cout % 32 == 0,
// at most 256 channels each side (above that the split into 32x32 blocks re-reads the activations too often; such
// layers are plain large GEMMs and stay on the library)
extern "C" int rss_conv_wgrad_supported(int Cin, int Cout, int ksize, int stride, int pad, int dil) {
    if (ksize != 1 && ksize != 3) return 0;
    if (stride != 1 && stride != 2) return 0;
    if (dil != 1 || pad != ksize / 2) return 0;
    if (Cin <= 0 || Cout <= 0 || Cin % kWgBlk || Cout % kWgBlk || Cin > 256 || Cout > 256) return 0;
    // a 1x1 conv has no tap reuse: with more than 4 (cout,cin) blocks the re-reads of X/dY cost more than the plain library GEMM
    if (ksize == 1 && (Cin / kWgBlk) * (Cout / kWgBlk) > 4) return 0;
    return 1;
}

// dw_acc (Cout,Cin,k,k) fp32 += weight gradient.  x (B,Hi,Wi,Cin), dy (B,Ho,Wo,Cout) bf16 NHWC; Ho = (Hi + 2p - k)/s + 1.
extern "C" int rss_conv_wgrad(const void* x, const void* dy, float* dw_acc, int B, int Hi, int Wi, int Cin, int Ho, int Wo, int Cout,
                              int ksize, int stride, int pad, int dil, cudaStream_t st) {
    if (!rss_conv_wgrad_supported(Cin, Cout, ksize, stride, pad, dil) || B <= 0 || Hi <= 0 || Wi <= 0) return RSS_ERR_SHAPE;
    if (Ho != (Hi + 2 * pad - ksize) / stride + 1 || Wo != (Wi + 2 * pad - ksize) / stride + 1) return RSS_ERR_SHAPE;
    if (((uintptr_t)dw_acc & 15) || ((uintptr_t)x & 15) || ((uintptr_t)dy & 15)) return RSS_ERR_SHAPE;
    WgGeom g;
    g.B = B; g.Hi = Hi; g.Wi = Wi; g.Cin = Cin; g.Ho = Ho; g.Wo = Wo; g.Cout = Cout; g.pad = pad; g.dil = dil;
    const int pix = stride == 1 ? 256 : 128;
    g.TW = Wo > 16 ? 32 : 16;
    g.TH = pix / g.TW;
    g.IH = (g.TH - 1) * stride + 1 + (ksize - 1) * dil;
    g.IW = (g.TW - 1) * stride + 1 + (ksize - 1) * dil;
    g.tiles_x = (Wo + g.TW - 1) / g.TW; g.tiles_y = (Ho + g.TH - 1) / g.TH;
    g.n_tiles = B * g.tiles_x * g.tiles_y;
    g.cin_blocks = Cin / kWgBlk;
    const int pairs = g.cin_blocks * (Cout / kWgBlk);
    const size_t stage = (size_t)(g.IH * g.IW + pix) * kWgPitch * 2;
    size_t smem = 2 * stage;
    const size_t flush = (size_t)kWgBlk * kWgBlk * ksize * ksize * sizeof(float);
    if (smem < flush) smem = flush;
    // split-K: enough CTAs to fill the machine twice over, never more than one tile per CTA
    int gx = (2 * num_sms() + pairs - 1) / pairs;
    if (gx > g.n_tiles) gx = g.n_tiles;
    if (gx < 1) gx = 1;
    dim3 grid(gx, pairs);
    cudaError_t e = cudaSuccess;
#define WG_LAUNCH(KS_, S_)                                                                                                    \
    do {                                                                                                                       \
        static size_t attr = 0;                                                                                                \
        if (smem > attr) {                                                                                                     \
            e = cudaFuncSetAttribute(conv_wgrad_kernel<KS_, S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(115 * 1024)); \
            attr = 115 * 1024;                                                                                                 \
        }                                                                                                                      \
        if (e == cudaSuccess)                                                                                                  \
            conv_wgrad_kernel<KS_, S_><<<grid, kWgThreads, smem, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw_acc, g); \
    } while (0)
    if (smem > 115 * 1024) return RSS_ERR_SHAPE;
    if (ksize == 3 && stride == 1) WG_LAUNCH(3, 1);
    else if (ksize == 3 && stride == 2) WG_LAUNCH(3, 2);
    else if (ksize == 1 && stride == 1) WG_LAUNCH(1, 1);
    else WG_LAUNCH(1, 2);
#undef WG_LAUNCH
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return RSS_ERR_CUDA; }
    return check_launch();
}
