// tcgen05 / TMA implicit-GEMM convolution for NHWC bf16 activations (sm_100a only).
//
// Covers the dense convolutions of the RSSFormer hot path that are stride-1 "same" convolutions:
//   * the FFN's three parallel convs dw(1x1) + dw6(3x3, dil 6) + dw12(3x3, dil 12) summed
//     (modules/ffn_block.py:226-228,250-257) as ONE GEMM: 17 taps (the three centre taps coincide and are
//     merged by adding their weights), K = 17 * 128;
//   * fc1 / fc2 (1x1, ffn_block.py:219,232), the neck's 480x480 1x1 (hrnet_aux.py:46), HRNet's 3x3/1x1
//     stride-1 convs (_hrnet_rssformer.py:209-213,253-259);
//   * their data gradients (same kernel, transposed weights, negated taps).
//
// GEMM view: M = B*H*W output pixels (BLOCK_M = 128 = BH rows x BW pixels of one image), N = Cout (BLOCK_N <= 256),
// K = taps x Cin in blocks of 64 channels.  Per k-block the A operand is ONE TMA tile load of the input at
// the tap-shifted coordinates {c0, x0+dx, y0+dy, b}: TMA's out-of-bounds zero fill IS the conv padding, so no
// im2col buffer and no boundary code exist.  Operands land in 128B-swizzled shared memory, tcgen05.mma
// (cta_group::1, M=128, kind::f16, bf16 x bf16 -> fp32) accumulates in TMEM, and a 4-warp epilogue drains
// TMEM with tcgen05.ld (+bias, ->bf16) while the MMA warp already works on the next tile (2 TMEM stages).
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue (two warps per TMEM lane quarter, each taking every
// other 16-column chunk: with the statistics epilogue four warps needed ~1.3x the MMA time of a tile).  Persistent: one CTA per SM.
#include <stdlib.h>
#include "tc05.cuh"

namespace rss {

constexpr int kBlockM = 128, kBlockK = 64, kUmmaK = 16;
constexpr int kConvThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two warps per TMEM lane quarter)
constexpr int kATileBytes = kBlockM * kBlockK * 2;          // 16 KB
constexpr int kMaxTaps = 32;

struct ConvTaps { int n; int dy[kMaxTaps]; int dx[kMaxTaps]; };

struct ConvGeom {
    int B, H, W, Cin, Cout;
    int BW, BH;                  // spatial tile: BW*BH == 128
    int tiles_x, tiles_y, m_tiles, n_tiles, block_n;
    int kchunks;                 // ceil(Cin / 64)
    int stages;
    int mm;                      // 128-row M sub-tiles per CTA tile (1 or 2): with 2, one B (weight) stage feeds two A tiles, i.e. the
                                 // weight traffic L2 -> shared memory per output pixel halves (the 17-tap FFN GEMM streams its
                                 // 557 KB of weights once per tile: ncu had the 1-sub-tile version pinned on the L2 -> SM fabric)
};

// K-major, 128B-swizzled operand tile (rows of 64 bf16 = 128 B; 8-row groups of 1024 B): cute::UMMA::SmemDescriptor
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);           // start address      bits [0,14)
    d |= (uint64_t)1 << 16;                                 // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                       // stride byte offset: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                                 // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                 // layout: SWIZZLE_128B
    return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=n
__host__ __device__ inline uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}

// Sum over the 32 lanes of a warp of 16 per-lane values at once ("transposed" butterfly: every exchange step halves the number of
// values a lane still carries, 8+4+2+1+1 = 16 shuffles instead of 16 x 5): returns the warp total of value index (lane >> 1) & 15.
__device__ __forceinline__ float warp_reduce16(float (&a)[16], int lane) {
#pragma unroll
    for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
        const bool hi = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const float send = hi ? a[i] : a[i + w];
            const float keep = hi ? a[i + w] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    return a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
}

// ---- the kernel -------------------------------------------------------------------------------------
// stat_accum != NULL: BatchNorm-statistics epilogue.  The epilogue warps also add, per output channel c, sum (y - K_c) and
// sum (y - K_c)^2 of the bf16-ROUNDED outputs into stat_accum[c] / stat_accum[Cout + c] (K = stat_shift - stat_shift_sub, NULL = 0): the "raw sums" of
// csrc/bn.cu's BnFin protocol, so the BatchNorm that follows needs no statistics pass over the tensor (FFN norm2 behind the
// dw + dw6 + dw12 GEMM: a 67 MB read per block and step).  Needs n_tiles == 1 and block_n <= 128.
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, ConvGeom g, ConvTaps taps,
                  float* __restrict__ stat_accum, const float* __restrict__ stat_shift, const float* __restrict__ stat_shift_sub) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t b_tile_bytes = (uint32_t)g.block_n * kBlockK * 2;
    const uint32_t a_bytes = (uint32_t)g.mm * kATileBytes;                 // mm sub-tiles of 128 rows, one TMA box (BH*mm image rows)
    const uint32_t stage_bytes = a_bytes + b_tile_bytes;                   // multiple of 1024 (block_n % 16 == 0 -> N*128 B % 2048)
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)g.stages * stage_bytes);
    // bars: [0,S) full, [S,2S) empty, [2S,2S+2) tmem_full, [2S+2,2S+4) tmem_empty ; then the TMEM base address
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * g.stages + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int acc_cols = g.mm * g.block_n;                                 // TMEM columns of one accumulator stage
    const uint32_t tmem_cols = (2 * acc_cols <= 32) ? 32 : (2 * acc_cols <= 64) ? 64 : (2 * acc_cols <= 128) ? 128
                               : (2 * acc_cols <= 256) ? 256 : 512;

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) { mbar_init(smem_u32(bars + s), 1); mbar_init(smem_u32(bars + g.stages + s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(bars + 2 * g.stages + s), 1); mbar_init(smem_u32(bars + 2 * g.stages + 2 + s), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_tiles = g.m_tiles * g.n_tiles;
    const int kblocks = taps.n * g.kchunks;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = tile / g.n_tiles, nt = tile % g.n_tiles;
                const int tx = mt % g.tiles_x, ty = (mt / g.tiles_x) % g.tiles_y, b = mt / (g.tiles_x * g.tiles_y);
                const int x0 = tx * g.BW, y0 = ty * g.BH * g.mm, n0 = nt * g.block_n;
                for (int kb = 0; kb < kblocks; ++kb) {
                    const int tap = kb / g.kchunks, kc = kb % g.kchunks;
                    mbar_wait(smem_u32(bars + g.stages + stage), phase ^ 1);           // slot free?
                    const uint32_t full = smem_u32(bars + stage);
                    const uint32_t a_dst = smem_u32(smem + (size_t)stage * stage_bytes);
                    mbar_expect_tx(full, stage_bytes);
                    tma_load_4d(a_dst, &tmap_a, full, kc * kBlockK, x0 + taps.dx[tap], y0 + taps.dy[tap], b);
                    tma_load_2d(a_dst + a_bytes, &tmap_b, full, kc * kBlockK, tap * g.Cout + n0);
                    if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: the whole warp runs the (warp-uniform) loop, one elected lane issues =================
        // (under `if (lane == 0)` the compiler cannot prove the descriptors warp-uniform and wraps every UTCHMMA in an
        //  ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall: ~150 cycles per MMA, tools/probe/mma_probe.cu)
        {
            const uint32_t idesc = make_idesc(g.block_n);
            const uint32_t leader = elect_one();
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint64_t desc_hi = make_sw128_desc(0);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(smem_u32(bars + 2 * g.stages + 2 + acc), acc_phase ^ 1);       // epilogue drained this TMEM stage?
                tc_fence_after();
                const uint32_t d_tmem = tmem_u + acc * acc_cols;
                for (int kb = 0; kb < kblocks; ++kb) {
                    const int kc = kb % g.kchunks;
                    int ksteps = (g.Cin - kc * kBlockK + kUmmaK - 1) / kUmmaK;            // skip zero-filled tail channels
                    if (ksteps > kBlockK / kUmmaK) ksteps = kBlockK / kUmmaK;
                    mbar_wait(smem_u32(bars + stage), phase);                             // TMA landed?
                    tc_fence_after();
                    const uint32_t a_lo = smem_u32(smem + (size_t)stage * stage_bytes) >> 4;
                    const uint32_t b_lo = a_lo + (a_bytes >> 4);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)     // +32 B (2 units of 16 B) per K=16 step inside the 128 B swizzle atom
                        if (k < ksteps) {
                            umma_bf16_elect(leader, d_tmem, desc_hi | (uint64_t)(a_lo + k * 2), desc_hi | (uint64_t)(b_lo + k * 2), idesc,
                                            (kb | k) != 0);
                            if (g.mm == 2)                         // second 128-row sub-tile: same weights, next 16 KB of A, next N columns
                                umma_bf16_elect(leader, d_tmem + g.block_n, desc_hi | (uint64_t)(a_lo + (kATileBytes >> 4) + k * 2),
                                                desc_hi | (uint64_t)(b_lo + k * 2), idesc, (kb | k) != 0);
                        }
                    umma_commit_elect(leader, smem_u32(bars + g.stages + stage));        // frees the smem slot when MMAs finish
                    if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit_elect(leader, smem_u32(bars + 2 * g.stages + acc));          // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue: TMEM -> registers -> (+bias, bf16) -> global (+ BatchNorm raw sums) =================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                // of the two warps of a quarter: takes the chunks with (chunk & 1) == half
        const int row = q * 32 + lane;                   // tile row == TMEM lane
        uint32_t acc = 0, acc_phase = 0;
        const bool stats = stat_accum != nullptr;
        float ssum[8], ssq[8];                           // per 16-channel chunk: this lane's channel is chunk*16 + (lane >> 1)
#pragma unroll
        for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / g.n_tiles, nt = tile % g.n_tiles;
            const int tx = mt % g.tiles_x, ty = (mt / g.tiles_x) % g.tiles_y, b = mt / (g.tiles_x * g.tiles_y);
            const int n0 = nt * g.block_n;
            mbar_wait(smem_u32(bars + 2 * g.stages + acc), acc_phase);
            tc_fence_after();
            if (!stats) {
                for (int m = 0; m < g.mm; ++m) {
                    const int x = tx * g.BW + row % g.BW, y = (ty * g.mm + m) * g.BH + row / g.BW;
                    const bool live = x < g.W && y < g.H;
                    __nv_bfloat16* dst = out + (((size_t)b * g.H + y) * g.W + x) * g.Cout + n0;
                    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * acc_cols + m * g.block_n;
                    for (int c0 = half * 16; c0 < g.block_n; c0 += 32) {
                        uint32_t r[16];
                        tmem_ld16(t_row + c0, r);
                        tmem_ld_wait();
                        if (live) {
                            float v[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + (bias ? bias[n0 + c0 + i] : 0.f);
                            store8(dst + c0, v);
                            store8(dst + c0 + 8, v + 8);
                        }
                    }
                }
            } else {
                // chunk-major order: both M sub-tiles of a 16-channel chunk are folded into one warp reduction
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const int c0 = ch * 16;
                    if ((ch & 1) == half && c0 < g.block_n) {
                        float bs[16], kk[16], s16[16], q16[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            bs[i] = bias ? bias[c0 + i] : 0.f;
                            kk[i] = (stat_shift ? stat_shift[c0 + i] : 0.f) - (stat_shift_sub ? stat_shift_sub[c0 + i] : 0.f);
                            s16[i] = 0.f; q16[i] = 0.f;
                        }
                        for (int m = 0; m < g.mm; ++m) {
                            const int x = tx * g.BW + row % g.BW, y = (ty * g.mm + m) * g.BH + row / g.BW;
                            const bool live = x < g.W && y < g.H;
                            __nv_bfloat16* dst = out + (((size_t)b * g.H + y) * g.W + x) * g.Cout;
                            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * acc_cols + m * g.block_n;
                            uint32_t r[16];
                            tmem_ld16(t_row + c0, r);
                            tmem_ld_wait();
                            if (live) {
                                float v[16];
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + bs[i];
                                store8(dst + c0, v);
                                store8(dst + c0 + 8, v + 8);
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    const float d = __bfloat162float(__float2bfloat16_rn(v[i])) - kk[i];
                                    s16[i] += d;
                                    q16[i] += d * d;
                                }
                            }
                        }
                        ssum[ch] += warp_reduce16(s16, lane);
                        ssq[ch] += warp_reduce16(q16, lane);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(bars + 2 * g.stages + 2 + acc));           // 8 warps -> count 8
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (stats) {
            // every MMA of this CTA has completed (the last accumulator was committed before its epilogue began), so the pipeline's
            // shared memory is free: the epilogue warps combine their channel totals there, then one atomic per channel and CTA
            float* red = reinterpret_cast<float*>(smem);                 // [4 quarters][2][128]; a (quarter, chunk) has one owner warp
            if ((lane & 1) == 0) {
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    if ((ch & 1) == half) {
                        red[(q * 2 + 0) * 128 + ch * 16 + (lane >> 1)] = ssum[ch];
                        red[(q * 2 + 1) * 128 + ch * 16 + (lane >> 1)] = ssq[ch];
                    }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");               // the 8 epilogue warps only
            const int c = (int)threadIdx.x - 64;
            if (c < g.block_n) {
                float s = 0.f, qq = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) { s += red[(w * 2 + 0) * 128 + c]; qq += red[(w * 2 + 1) * 128 + c]; }
                atomicAdd(stat_accum + c, s);
                atomicAdd(stat_accum + g.Cout + c, qq);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ---- weight packing: fp32 (Cout,Cin,k,k) sources -> bf16 [tap][N][K] (K contiguous) ------------------------
struct PackSrc { const float* w; const float* b; int k; int dil; };
struct PackPlan {
    int n_src;
    PackSrc src[3];
    int n_taps;
    int tap_src[kMaxTaps][3];        // per tap: up to 3 contributing (source, ky*k+kx); -1 = none
    int tap_pos[kMaxTaps][3];
};

// transpose == 0: packed[tap][co][ci] = sum_s w_s[co][ci][pos]      (forward:  N=Cout, K=Cin)
// transpose == 1: packed[tap][ci][co] = sum_s w_s[co][ci][pos]      (dgrad:    N=Cin,  K=Cout)
__global__ void conv_pack_kernel(PackPlan plan, int Cout, int Cin, int transpose, __nv_bfloat16* __restrict__ packed,
                                 float* __restrict__ bias_sum) {
    const int64_t per_tap = (int64_t)Cout * Cin;
    const int64_t total = per_tap * plan.n_taps;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int tap = (int)(idx / per_tap);
        const int64_t r = idx % per_tap;
        const int n = (int)(r / (transpose ? Cout : Cin)), k = (int)(r % (transpose ? Cout : Cin));
        const int co = transpose ? k : n, ci = transpose ? n : k;
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const int s = plan.tap_src[tap][e];
            if (s >= 0) {
                const int kk = plan.src[s].k * plan.src[s].k;
                acc += plan.src[s].w[((int64_t)co * Cin + ci) * kk + plan.tap_pos[tap][e]];
            }
        }
        packed[idx] = __float2bfloat16_rn(acc);
    }
    if (bias_sum && blockIdx.x == 0) {
        for (int c = threadIdx.x; c < Cout; c += blockDim.x) {
            float s = 0.f;
            bool any = false;
            for (int e = 0; e < plan.n_src; ++e) if (plan.src[e].b) { s += plan.src[e].b[c]; any = true; }
            bias_sum[c] = any ? s : 0.f;
        }
    }
}

static int build_plan(const float* const* w, const float* const* b, const int* ks, const int* dils, int n_src, int negate,
                      PackPlan* plan, ConvTaps* taps) {
    if (n_src < 1 || n_src > 3) return RSS_ERR_SHAPE;
    plan->n_src = n_src;
    taps->n = 0;
    for (int t = 0; t < kMaxTaps; ++t) for (int e = 0; e < 3; ++e) { plan->tap_src[t][e] = -1; plan->tap_pos[t][e] = 0; }
    for (int s = 0; s < n_src; ++s) {
        const int k = ks[s], d = dils[s];
        if (k != 1 && k != 3) return RSS_ERR_SHAPE;
        plan->src[s].w = w[s]; plan->src[s].b = b ? b[s] : nullptr; plan->src[s].k = k; plan->src[s].dil = d;
        for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx) {
                int dy = (ky - k / 2) * d, dx = (kx - k / 2) * d;
                if (negate) { dy = -dy; dx = -dx; }
                int t = 0;
                for (; t < taps->n; ++t) if (taps->dy[t] == dy && taps->dx[t] == dx) break;
                if (t == taps->n) {
                    if (taps->n == kMaxTaps) return RSS_ERR_SHAPE;
                    taps->dy[t] = dy; taps->dx[t] = dx; ++taps->n;
                }
                int e = 0;
                while (e < 3 && plan->tap_src[t][e] >= 0) ++e;
                if (e == 3) return RSS_ERR_SHAPE;
                plan->tap_src[t][e] = s; plan->tap_pos[t][e] = ky * k + kx;
            }
    }
    plan->n_taps = taps->n;
    return RSS_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the driver entry point is resolved through the runtime (no link-time dependency on libcuda.so)
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int make_maps(const void* x, const void* wp, const ConvGeom& g, int n_taps, CUtensorMap* ma, CUtensorMap* mb) {
    EncodeTiledFn cuTensorMapEncodeTiled = encode_tiled();
    if (!cuTensorMapEncodeTiled) return RSS_ERR_CUDA;
    {
        cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.B};
        cuuint64_t strides[3] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.W * g.Cin * 2, (cuuint64_t)g.H * g.W * g.Cin * 2};
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)g.BW, (cuuint32_t)(g.BH * g.mm), 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = cuTensorMapEncodeTiled(ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return RSS_ERR_CUDA; }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)g.Cin, (cuuint64_t)n_taps * g.Cout};
        cuuint64_t strides[1] = {(cuuint64_t)g.Cin * 2};
        cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)g.block_n};
        cuuint32_t es[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), dims, strides, box, es,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return RSS_ERR_CUDA; }
    }
    return RSS_OK;
}

}  // namespace rss

using namespace rss;

// geometry the igemm path accepts: stride 1, "same" padding, Cin % 8 == 0, Cout % 16 == 0, W a power of two <= 128 or a
// multiple of 128; returns 1 if supported
extern "C" int rss_conv_igemm_supported(int B, int H, int W, int Cin, int Cout) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin < 16 || Cin % 8 || Cout < 16 || Cout % 16) return 0;
    if (W >= 128) return W % 128 == 0;
    return (W & (W - 1)) == 0 && W >= 8;
}

extern "C" size_t rss_conv_packed_bytes(int n_srcs, const int* ksizes, int Cout, int Cin) {
    size_t taps = 0;
    for (int s = 0; s < n_srcs; ++s) taps += (size_t)ksizes[s] * ksizes[s];
    return taps * Cout * Cin * 2;     // upper bound (merged centre taps make it smaller)
}

// Pack 1..3 parallel convolutions (same Cout/Cin, summed outputs) into the igemm weight layout.
// transpose=0: forward operand; transpose=1: data-gradient operand (taps negated, Cout/Cin swapped).
extern "C" int rss_conv_pack_weights(const float* const* weights, const float* const* biases, const int* ksizes, const int* dilations,
                                     int n_srcs, int Cout, int Cin, int transpose, void* packed, float* bias_sum,
                                     int* n_taps_out, int* taps_dy_out, int* taps_dx_out, cudaStream_t st) {
    PackPlan plan; ConvTaps taps;
    int rc = build_plan(weights, biases, ksizes, dilations, n_srcs, transpose, &plan, &taps);
    if (rc != RSS_OK) return rc;
    const int64_t total = (int64_t)plan.n_taps * Cout * Cin;
    int grid = (int)((total + 255) / 256);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    conv_pack_kernel<<<grid, 256, 0, st>>>(plan, Cout, Cin, transpose, (__nv_bfloat16*)packed, transpose ? nullptr : bias_sum);
    *n_taps_out = taps.n;
    for (int t = 0; t < taps.n; ++t) { taps_dy_out[t] = taps.dy[t]; taps_dx_out[t] = taps.dx[t]; }
    return check_launch();
}

// y[b,y,x,:] = bias + sum_tap W_tap . x[b, y+dy_tap, x+dx_tap, :]   (zero outside the image); bf16 in/out, fp32 accumulate.
// w_packed: bf16 [n_taps][Cout][Cin].  For a data gradient call it with (x=dY, Cin<->Cout swapped, transposed pack).
// stat_accum / stat_shift: BatchNorm-statistics epilogue (see conv_igemm_kernel); Cout <= 128 only.
extern "C" int rss_conv_igemm_stats(const void* x, const void* w_packed, const float* bias, void* y, int B, int H, int W, int Cin, int Cout,
                                    int n_taps, const int* taps_dy, const int* taps_dx, float* stat_accum, const float* stat_shift,
                                    const float* stat_shift_sub, cudaStream_t st) {
    if (!rss_conv_igemm_supported(B, H, W, Cin, Cout) || n_taps < 1 || n_taps > kMaxTaps) return RSS_ERR_SHAPE;
    if (stat_accum && Cout > 128) return RSS_ERR_SHAPE;
    ConvGeom g;
    g.B = B; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout;
    g.BW = W >= 128 ? 128 : W; g.BH = kBlockM / g.BW;
    g.n_tiles = (Cout + 255) / 256;
    if (Cout % g.n_tiles || (Cout / g.n_tiles) % 16) return RSS_ERR_SHAPE;
    g.block_n = Cout / g.n_tiles;
    // two M sub-tiles per CTA tile when both accumulator stages still fit in TMEM (2 * 2 * N <= 512 columns), the image rows pair
    // up, and the launch still has at least 64 tiles; RSS_IGEMM_MM=1 keeps the single sub-tile version (A/B measurements)
    static const int mm_env = getenv("RSS_IGEMM_MM") ? atoi(getenv("RSS_IGEMM_MM")) : 2;
    g.mm = 1;
    if (mm_env >= 2 && g.block_n <= 128 && g.BH * 2 <= 256 && H % (2 * g.BH) == 0 &&
        (int64_t)B * ((W + g.BW - 1) / g.BW) * (H / (2 * g.BH)) * g.n_tiles >= 64)
        g.mm = 2;
    g.tiles_x = (W + g.BW - 1) / g.BW; g.tiles_y = (H + g.BH * g.mm - 1) / (g.BH * g.mm);
    g.m_tiles = B * g.tiles_x * g.tiles_y;
    g.kchunks = (Cin + kBlockK - 1) / kBlockK;
    const size_t stage_bytes = (size_t)g.mm * kATileBytes + (size_t)g.block_n * kBlockK * 2;
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return RSS_ERR_SHAPE;
    g.stages = stages;
    const size_t smem = 1024 + stages * stage_bytes + (2 * stages + 4) * 8 + 16;
    ConvTaps taps; taps.n = n_taps;
    for (int t = 0; t < n_taps; ++t) { taps.dy[t] = taps_dy[t]; taps.dx[t] = taps_dx[t]; }
    CUtensorMap ma, mb;
    int rc = make_maps(x, w_packed, g, n_taps, &ma, &mb);
    if (rc != RSS_OK) return rc;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        if (e != cudaSuccess) { g_last_cuda_error = (int)e; return RSS_ERR_CUDA; }
        attr_smem = 227 * 1024;
    }
    int grid = num_sms();
    const int total_tiles = g.m_tiles * g.n_tiles;
    if (grid > total_tiles) grid = total_tiles;
    conv_igemm_kernel<<<grid, kConvThreads, smem, st>>>(ma, mb, bias, (__nv_bfloat16*)y, g, taps, stat_accum, stat_shift, stat_shift_sub);
    return check_launch();
}

extern "C" int rss_conv_igemm(const void* x, const void* w_packed, const float* bias, void* y, int B, int H, int W, int Cin, int Cout,
                              int n_taps, const int* taps_dy, const int* taps_dx, cudaStream_t st) {
    return rss_conv_igemm_stats(x, w_packed, bias, y, B, H, W, Cin, Cout, n_taps, taps_dy, taps_dx, nullptr, nullptr, nullptr, st);
}
