// tcgen05 weight gradient of the stride-1 HRNet-family convolutions (3x3 pad 1 and 1x1; _hrnet_rssformer.py:209-287):
//
//   dW[tap][co][ci] += sum over padded-linear positions q of  dY[q][co] * T(X)[q + dy_tap*(W+2) + dx_tap][ci]
//
// Same staging as conv_cf.cu: TMA box loads of whole zero-padded rows ({64 ch, W+2 pixels from x=-1, NR rows}; out-of-bounds
// zero fill = padding, so dY is zero on the padding columns and X is zero outside the image) land as one 128-byte, 128B-swizzled
// row per position.  With positions as the K dimension that layout is exactly the canonical MN-major SWIZZLE_128B UMMA operand
// (64 contiguous channels per row, rows = K index, 8-row swizzle period), for BOTH operands: A = dY^T (M = cout), B = X^T (N = cin).
// A tap is a shift of the B start address by whole rows (the swizzle XOR uses absolute address bits, see conv_cf.cu).
// Each MMA is M=128 (cout rows beyond Cout read don't-care bytes and land in unused TMEM lanes) x N=cin block x K=16 positions;
// every tap owns N TMEM columns and accumulates over the CTA's whole share of the positions (split-K over CTAs); at the end the
// accumulators are staged through shared memory into the PyTorch (Cout,Cin,k,k) order and added to the caller's fp32 gradient
// with 16-byte vector reductions.  Optional: the previous layer's BN+ReLU applied to X in place after it lands (helper warps).
#include <stdlib.h>
#include "tc05.cuh"

namespace rss {

constexpr int kWtThreads = 288;          // warp 0 TMA producer, warps 1-3 transform helpers, warp 4 MMA issuer, warps 5-8 epilogue
constexpr int kWtHelpers = 96;
constexpr int kWtKT = 128;               // positions per pipeline stage (8 K=16 MMA steps)

struct WtGeom {
    int B, H, W, Cin, Cout;
    int halo, Wp, Q;
    int tiles_per_img, n_tiles;
    int NRx, Px, NRd, Pd;                // staged rows / plane pitch (positions, multiple of 8) of X and dY
    int KCx, KCd;                        // 64-channel planes staged per tile (of this CTA's cin / cout block)
    int ci_blk, co_blk;                  // channels per CTA block (<= 128)
    int n_ci_blk, n_co_blk, n_tap_grp, taps_per_grp;
    int n_taps;
    int tap_off[9];
    int ks;                              // kernel size (for the output index)
    int in_relu;
};

__device__ __forceinline__ uint64_t wt_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {     // MN-major, SWIZZLE_128B
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;       // next 64-element M/N group (= next channel plane)
    d |= (uint64_t)(1024 >> 4) << 32;                       // next 8 K rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t wt_idesc(int n) {       // kind::f16, D=f32, A=B=bf16, both MN-major, M=128
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ int wt_row_lo(int q0, int halo, int Wp) { return (q0 - halo + Wp) / Wp - 1 - halo; }
__device__ __forceinline__ void wt_red4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kWtThreads, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                     float* __restrict__ dw, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                     const __grid_constant__ WtGeom g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t x_bytes = (uint32_t)g.KCx * g.Px * 128, d_bytes = (uint32_t)g.KCd * g.Pd * 128;
    const uint32_t stage_bytes = x_bytes + d_bytes;
    const uint32_t s0 = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * (size_t)stage_bytes);
    uint64_t* bar_landed = bars, *bar_ready = bars + 2, *bar_empty = bars + 4, *bar_done = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool xform = in_scale != nullptr;
    // this CTA's block of the gradient
    int by = blockIdx.y;
    const int tg = by % g.n_tap_grp; by /= g.n_tap_grp;
    const int cib = by % g.n_ci_blk, cob = by / g.n_ci_blk;
    const int tap0 = tg * g.taps_per_grp;
    const int ntap = (g.n_taps - tap0 < g.taps_per_grp) ? g.n_taps - tap0 : g.taps_per_grp;
    const int ci0 = cib * g.ci_blk, co0 = cob * g.co_blk;
    const int N = g.ci_blk;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < ntap * N) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(bar_landed + s), 1); mbar_init(smem_u32(bar_ready + s), kWtHelpers / 32); mbar_init(smem_u32(bar_empty + s), 1);
        }
        mbar_init(smem_u32(bar_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_dy) : "memory");
    }
    if (warp == 4) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)(g.KCx * g.NRx + g.KCd * g.NRd) * g.Wp * 128;
            int i = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++i) {
                const int si = i & 1, use = i >> 1;
                if (use > 0) mbar_wait(smem_u32(bar_empty + si), (use - 1) & 1);
                const int b = tile / g.tiles_per_img, t = tile % g.tiles_per_img;
                const int q0 = t * kWtKT;
                const int rx = wt_row_lo(q0, g.halo, g.Wp), rd = q0 / g.Wp;
                const uint32_t full = smem_u32(bar_landed + si);
                const uint32_t ds = s0 + si * stage_bytes, xs = ds + d_bytes;
                mbar_expect_tx(full, tx_bytes);
                for (int pl = 0; pl < g.KCx; ++pl)
                    tma_load_4d(xs + (uint32_t)(pl * g.Px) * 128, &tmap_x, full, ci0 + pl * 64, -g.halo, rx, b);
                for (int pl = 0; pl < g.KCd; ++pl)
                    tma_load_4d(ds + (uint32_t)(pl * g.Pd) * 128, &tmap_dy, full, co0 + pl * 64, -g.halo, rd, b);
            }
        }
    } else if (warp < 4) {
        // ================= helpers: BN(+ReLU) of the previous layer applied to the landed X rows =================
        if (xform) {
            const int CHb = (g.ci_blk < 64 ? g.ci_blk : 64) >> 3;              // 16-byte chunks per row that hold real channels
            const int nchunk = g.KCx * CHb;                                     // chunks per position over all planes
            const int ht = threadIdx.x - 32;
            const int npos = g.NRx * g.Wp;
            int i = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++i) {
                const int si = i & 1;
                mbar_wait(smem_u32(bar_landed + si), (i >> 1) & 1);
                const int t = tile % g.tiles_per_img;
                const int rx = wt_row_lo(t * kWtKT, g.halo, g.Wp);
                uint8_t* base = smem + (size_t)si * stage_bytes + d_bytes;
                for (int e = ht; e < npos * nchunk; e += kWtHelpers) {
                    const int p = e / nchunk, cc = e % nchunk;
                    const int pl = cc / CHb, ch = cc % CHb;
                    const int r = rx + p / g.Wp, c = p % g.Wp - g.halo;
                    if (r >= 0 && r < g.H && c >= 0 && c < g.W) {
                        uint4* ptr = reinterpret_cast<uint4*>(base + ((size_t)pl * g.Px + p) * 128 + ((ch ^ (p & 7)) << 4));
                        Raw8<__nv_bfloat16> raw;
                        raw.r = *ptr;
                        float v[8];
                        unpack8(raw, v);
                        const int cg = ci0 + pl * 64 + ch * 8;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            v[k] = fmaf(v[k], __ldg(in_scale + cg + k), __ldg(in_shift + cg + k));
                            if (g.in_relu) v[k] = fmaxf(v[k], 0.f);
                        }
                        store8(reinterpret_cast<__nv_bfloat16*>(ptr), v);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(bar_ready + si));
            }
        }
    } else if (warp == 4) {
        // ================= MMA issuer: the whole warp runs the (warp-uniform) loop, one elected lane issues =================
        {
            const uint32_t idesc = wt_idesc(N);
            uint64_t* bar_in = xform ? bar_ready : bar_landed;
            const uint64_t a_hi = wt_desc_mn(0, (uint32_t)g.Pd * 128), b_hi = wt_desc_mn(0, (uint32_t)g.Px * 128);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t leader = elect_one();
            int i = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++i) {
                const int si = i & 1;
                mbar_wait(smem_u32(bar_in + si), (i >> 1) & 1);
                fence_proxy_async_smem();
                tc_fence_after();
                const int t = tile % g.tiles_per_img;
                const int q0 = t * kWtKT;
                const int px0 = q0 - wt_row_lo(q0, g.halo, g.Wp) * g.Wp;       // staged X index of position q0 (tap 0,0)
                const int pd0 = q0 - (q0 / g.Wp) * g.Wp;                        // staged dY index of position q0
                const uint32_t ds = s0 + si * stage_bytes, xs = ds + d_bytes;
                // only the 14-bit start-address field (address >> 4; one 128-byte row = 8 units) changes between MMAs
                uint32_t a_lo = (ds >> 4) + (uint32_t)pd0 * 8;
                uint32_t b_lo = (xs >> 4) + (uint32_t)px0 * 8;
                for (int ks = 0; ks < kWtKT / 16; ++ks) {
                    const uint64_t adesc = a_hi | (uint64_t)a_lo;
                    const uint32_t accum = (i | ks) != 0;
                    for (int tp = 0; tp < ntap; ++tp)
                        umma_bf16_elect(leader, tmem_u + tp * N, adesc, b_hi | (uint64_t)(b_lo + (uint32_t)(g.tap_off[tap0 + tp] * 8)), idesc, accum);
                    a_lo += 16 * 8;
                    b_lo += 16 * 8;
                }
                umma_commit_elect(leader, smem_u32(bar_empty + si));
            }
            umma_commit_elect(leader, smem_u32(bar_done));
        }
    }
    // ================= flush: TMEM -> smem [co][ci][tap] -> vector reductions into dW =================
    if (warp >= 5) {
        mbar_wait(smem_u32(bar_done), 0);
        tc_fence_after();
    }
    __syncthreads();                                            // all MMAs retired: the staging ring is free
    float* stg = reinterpret_cast<float*>(smem);
    const int T = g.ks * g.ks;                                  // taps in the output layout (== n_taps == ntap: one tap group)
    const int row = N * ntap;                                   // floats per cout row of this block, contiguous in dW
    const int pitch = row + 4;                                  // +16 B: rows of consecutive lanes fall into different banks
    if (warp >= 5) {
        const int q4 = warp & 3;
        const int m = q4 * 32 + lane;                           // TMEM lane == cout row of the block
        const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16);
        for (int tp = 0; tp < ntap; ++tp)
            for (int c0 = 0; c0 < N; c0 += 16) {
                uint32_t rr[16];
                tmem_ld16(t_row + tp * N + c0, rr);
                tmem_ld_wait();
                if (m < g.co_blk) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) stg[(size_t)m * pitch + (c0 + k) * ntap + tp] = __uint_as_float(rr[k]);
                }
            }
    }
    tc_fence_before();
    __syncthreads();
    if ((row & 3) == 0 && g.n_tap_grp == 1) {
        for (int e = threadIdx.x * 4; e < g.co_blk * row; e += kWtThreads * 4) {
            const int m = e / row, off = e % row;
            const float* src = stg + (size_t)m * pitch + off;
            wt_red4(dw + ((size_t)(co0 + m) * g.Cin + ci0) * T + off, src[0], src[1], src[2], src[3]);
        }
    } else {
        for (int e = threadIdx.x; e < g.co_blk * row; e += kWtThreads) {
            const int m = e / row, off = e % row;
            const int ci = off / ntap, tp = off % ntap;
            atomicAdd(dw + ((size_t)(co0 + m) * g.Cin + ci0 + ci) * T + tap0 + tp, stg[(size_t)m * pitch + off]);
        }
    }
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, tmem_cols);
}

struct WtPlan { WtGeom g; size_t smem; dim3 grid; };

static int wt_plan(int B, int H, int W, int Cin, int Cout, int ksize, WtPlan* pl) {
    if (B <= 0 || H <= 0 || W <= 0 || (ksize != 1 && ksize != 3)) return RSS_ERR_SHAPE;
    if (Cin % 32 || Cout % 32 || Cin < 32 || Cout < 32 || Cin > 512 || Cout > 512) return RSS_ERR_SHAPE;
    WtGeom& g = pl->g;
    g.B = B; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.ks = ksize; g.in_relu = 0;
    g.halo = ksize / 2; g.Wp = W + 2 * g.halo; g.Q = H * g.Wp;
    if (g.Wp > 256) return RSS_ERR_SHAPE;
    g.n_taps = ksize * ksize;
    for (int a = 0, n = 0; a < ksize; ++a) for (int b = 0; b < ksize; ++b, ++n) g.tap_off[n] = (a - g.halo) * g.Wp + (b - g.halo);
    // every CTA owns ALL taps of a (<=128 cout) x (32 cin) block: 9 x 32 = 288 TMEM columns, and each cout row of the block is one
    // contiguous run of 32*k*k floats in the (Cout,Cin,k,k) gradient -> 16-byte vector reductions.  (A first version gave each CTA
    // 128 cin and a third of the taps: 3-float runs, scalar atomics, 130 us on the 256-channel layers.)
    g.ci_blk = 32; g.co_blk = Cout > 128 ? 128 : Cout;
    if (Cin % g.ci_blk || Cout % g.co_blk) return RSS_ERR_SHAPE;
    if (g.ci_blk % 16) return RSS_ERR_SHAPE;
    g.n_ci_blk = Cin / g.ci_blk; g.n_co_blk = Cout / g.co_blk;
    int tpg = 512 / g.ci_blk;                                  // taps whose accumulators fit in the 512 TMEM columns
    if (tpg > g.n_taps) tpg = g.n_taps;
    g.n_tap_grp = (g.n_taps + tpg - 1) / tpg;
    g.taps_per_grp = (g.n_taps + g.n_tap_grp - 1) / g.n_tap_grp;
    g.KCx = (g.ci_blk + 63) / 64; g.KCd = (g.co_blk + 63) / 64;
    const int L = kWtKT + 2 * g.halo;
    g.NRx = (L + g.Wp - 2) / g.Wp + 1 + 2 * g.halo;
    g.NRd = (kWtKT + g.Wp - 2) / g.Wp + 1;
    if (g.NRx > 256) return RSS_ERR_SHAPE;
    g.Px = (g.NRx * g.Wp + 7) & ~7; g.Pd = (g.NRd * g.Wp + 7) & ~7;
    const size_t stage = ((size_t)g.KCx * g.Px + (size_t)g.KCd * g.Pd) * 128;
    // the M=128 A operand always reads two 64-channel groups (LBO = one dY plane apart): each stage holds the dY planes first, so
    // with a single dY plane the second (don't-care) group falls into the X planes of the same stage
    size_t need = 2 * stage + 16 * 8 + 64;
    if ((size_t)g.KCd * g.Pd + (size_t)g.KCx * g.Px < (size_t)2 * g.Pd) return RSS_ERR_SHAPE;
    const size_t flush = (size_t)g.co_blk * (g.ci_blk * g.taps_per_grp + 4) * 4;
    if (flush > need) need = flush;
    if (need + 1024 > 225 * 1024) return RSS_ERR_SHAPE;
    pl->smem = need + 1024;
    g.tiles_per_img = (g.Q + kWtKT - 1) / kWtKT;
    g.n_tiles = B * g.tiles_per_img;
    const int gy = g.n_tap_grp * g.n_ci_blk * g.n_co_blk;
    int gx = (num_sms() + gy - 1) / gy;
    if (gx > g.n_tiles) gx = g.n_tiles;
    if (gx < 1) gx = 1;
    pl->grid = dim3(gx, gy);
    return RSS_OK;
}

typedef CUresult (*WtEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static WtEncodeTiledFn wt_encode_tiled() {
    static WtEncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (WtEncodeTiledFn)p;
    }
    return fn;
}
static int wt_map(CUtensorMap* tm, const void* base, int C, int W, int H, int B, int box_w, int box_rows) {
    WtEncodeTiledFn enc = wt_encode_tiled();
    if (!enc) return RSS_ERR_CUDA;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_rows, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return RSS_ERR_CUDA; }
    return RSS_OK;
}

}  // namespace rss

using namespace rss;

extern "C" int rss_conv_wgrad_tc_supported(int B, int H, int W, int Cin, int Cout, int ksize) {
    WtPlan pl;
    return wt_plan(B, H, W, Cin, Cout, ksize, &pl) == RSS_OK;
}

// dw_acc (Cout,Cin,k,k) fp32 += weight gradient of a stride-1 "same" k x k conv (k in {1,3}); x (B,H,W,Cin), dy (B,H,W,Cout) bf16
// NHWC.  in_scale/in_shift/in_relu: optional per-channel transform relu(x*scale+shift) applied to x on load (the conv consumed the
// BatchNorm+ReLU of a tensor that was never materialised).
extern "C" int rss_conv_wgrad_tc(const void* x, const void* dy, float* dw_acc, int B, int H, int W, int Cin, int Cout, int ksize,
                                 const float* in_scale, const float* in_shift, int in_relu, cudaStream_t stream) {
    WtPlan pl;
    int rc = wt_plan(B, H, W, Cin, Cout, ksize, &pl);
    if (rc != RSS_OK) return rc;
    if (((uintptr_t)dw_acc & 15) || ((uintptr_t)x & 15) || ((uintptr_t)dy & 15)) return RSS_ERR_SHAPE;
    if ((in_scale == nullptr) != (in_shift == nullptr)) return RSS_ERR_SHAPE;
    pl.g.in_relu = in_relu;
    CUtensorMap tx, td;
    rc = wt_map(&tx, x, Cin, W, H, B, pl.g.Wp, pl.g.NRx);
    if (rc != RSS_OK) return rc;
    rc = wt_map(&td, dy, Cout, W, H, B, pl.g.Wp, pl.g.NRd);
    if (rc != RSS_OK) return rc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(225 * 1024));
        if (e != cudaSuccess) { g_last_cuda_error = (int)e; (void)cudaGetLastError(); return RSS_ERR_CUDA; }
        attr = true;
    }
    conv_wgrad_tc_kernel<<<pl.grid, kWtThreads, pl.smem, stream>>>(tx, td, dw_acc, in_scale, in_shift, pl.g);
    return check_launch();
}
