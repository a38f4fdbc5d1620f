// Fused tcgen05 convolution for the HRNet family (BasicBlock/Bottleneck 3x3 and 1x1 stride-1 convs and their data
// gradients): _hrnet_rssformer.py:209-287.  NHWC bf16 in/out, fp32 TMEM accumulation.
//
//     y = E( conv( T(x) ) [+ add] )
//       T = identity | relu(x*scale + shift)        the BatchNorm(+ReLU) of the PREVIOUS layer, applied to the staged tile
//       E = identity
//         | statistics:  per-channel batch statistics of y (the BatchNorm that FOLLOWS), finalised by the last CTA
//         | bn-backward: y is the gradient w.r.t. the OUTPUT of a BatchNorm(+ReLU) whose input z is given: the kernel stores
//                        g = y * relu_mask and produces the two reductions of the BatchNorm backward, sum(g) and sum(g * xhat)
//
// Why not the tap-shifted-TMA kernel of conv_igemm.cu: that one re-fetches the A tile from L2 once per tap (9x for a 3x3).
// Here every input pixel is staged in shared memory ONCE per tile and all taps read it at shifted addresses:
//   * positions are linearised over the zero-padded image, q = row*(W+2) + col+1, so that tap (dy,dx) of output q is input
//     q + dy*(W+2) + dx: a pure address offset (outputs that fall on a padding column are computed and dropped, 2/(W+2) waste);
//   * the staged tile is one 128-byte row (64 channels) per position in the canonical 128B-swizzled K-major UMMA layout
//     (16-byte chunk c of position p lives at p*128 + ((c ^ (p & 7)) << 4)), so a tap shift is "start address += offset*128 B"
//     in the shared-memory descriptor.  Measured on B200: the hardware applies the swizzle XOR to the ABSOLUTE shared-memory
//     address bits [7,10), so a start address on any 128-byte row works with base_offset = 0;
//   * the tile is staged by ONE TMA box load per 64-channel plane: whole padded rows {64 ch, W+2 pixels from x=-1, NR rows from
//     y=r_lo} -- TMA's out-of-bounds zero fill IS the padding, and the dense box order [row][x][64 ch] IS the padded-linear
//     position order.
//
// Version 2 (round 2).  The round-1 kernel was bit-correct but took 37 us on the branch-0 layer (33.5 MB, HBM floor 5 us); a
// clock64() trace of one CTA (tools/cf_trace.py, gpurun_out/cf_trace.json) showed where: the MMA issue loop cost ~230 cycles per
// TAP (dynamically indexed constant loads + R2UR moves per descriptor: 2400 cycles per 128-row block whose tensor-pipe time is
// 288), the 96 helper threads needed 6000 cycles per tile for the input transform, and the 4 epilogue warps 650 cycles per
// block.  Now:
//   * Cin, Cout and the tap count are template parameters: the issue loop is straight-line code whose descriptors differ by
//     compile-time constants from two uniform registers;
//   * 8 epilogue warps: both groups drain the SAME accumulator, each thread one pixel x Cout/2 channels; the operands of the
//     fused epilogues (residual, z) are prefetched before the accumulator is ready and TMEM is released right after tcgen05.ld;
//   * 4 transform warps walk the interior of the staged rows (padding stays exactly zero) with a fixed channel chunk per thread;
//   * two 128-row blocks per tile whenever shared memory and TMEM allow (less halo re-fetch).
#include <stdlib.h>
#include "tc05.cuh"

namespace rss {

// warp 0 TMA producer, warps 1 and 15 MMA issuers (even / odd 128-row blocks), warps 2-5 input transform, warps 6-13 epilogue,
// warp 14 copy warp (bulk stores/loads)
constexpr int kCfThreads = 512;
constexpr int kCfXfThreads = 128;
constexpr int kCfEpiThreads = 256;
constexpr int kCfMaxTaps = 9;
constexpr int kCfMaxStages = 3;
constexpr int kCfPlain = 0, kCfStats = 1, kCfBnRed = 2;

struct CfGeom {
    int B, H, W;
    int halo, Wp, Q;                     // padded pitch W + 2*halo, positions per image H*Wp
    int MM, MT;                          // 128-row MMA blocks per tile, MT = 128*MM
    int tiles_per_img, n_tiles;
    int NR, P;                           // staged padded rows per tile; plane pitch in positions (>= NR*Wp, multiple of 16)
    int S;                               // ring stages (2 or 3)
    int tap_off[kCfMaxTaps];             // dy*Wp + dx (signed)
    int in_relu;
    int xsplit, box_w;                   // Wp > 256 (TMA box limit): every staged row is fetched as xsplit boxes of box_w positions
    int w_row_stride, w_tap_stride;      // weight operand: element (tap, n, k) at n*w_row_stride + tap*w_tap_stride + k (k contiguous)
    unsigned int wp_magic;               // ceil(2^32 / Wp): q / Wp == __umulhi(q, wp_magic) for q < 2^32 / Wp (positions are < 2^17)
    int dbg;                             // profiling aid (RSS_CF_DBG bits): 1 epilogue does no math / staging stores, 4 no MMAs issued
    int ops_staged;                      // epilogue operands (add / bn_z / bn_out) are staged in shared memory by the copy warp
    long long* trace;                    // profiling aid (RSS_CF_TRACE_PTR): CTA 0 records clock64() per role/tile/event, [4 roles][16 tiles][8]
};
#ifdef RSS_CF_TRACE_BUILD
#define CF_TRACE(role, i, k) do { if (g.trace && blockIdx.x == 0 && (i) < 16) g.trace[((role) * 16 + (i)) * 8 + (k)] = clock64(); } while (0)
#else
#define CF_TRACE(role, i, k) do { } while (0)
#endif

struct CfEpi {
    const __nv_bfloat16* add;            // [B,H,W,Cout] added to the accumulator before anything else (NULL: none)
    // shared by the two reducing epilogues
    float* accum;                        // [2*Cout] persistent, zero between launches
    unsigned int* ticket;                // persistent, zero between launches
    float count;                         // B*H*W
    // kCfStats
    const float* gamma; const float* beta;
    float* running_mean; float* running_var;     // may be NULL
    float momentum, eps;
    float* mean_out; float* invstd_out; float* scale_out; float* shift_out;
    // kCfBnRed
    const __nv_bfloat16* bn_z;           // [B,H,W,Cout] input of the BatchNorm being differentiated
    const __nv_bfloat16* bn_out;         // [B,H,W,Cout] its activated output (mask = out > 0) or NULL (mask from z and the affine)
    const float* bn_mean; const float* bn_invstd; const float* bn_scale; const float* bn_shift;
    int bn_relu;
    float* sums_out;                     // [2*Cout]: sum(g), sum(g*xhat)
};

__host__ __device__ constexpr uint32_t cf_idesc(int n) {    // kind::f16, D=f32, A=B=bf16, K-major both, M=128
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// K-major swizzled operand descriptor without the start address: ROWB = 128 -> SWIZZLE_128B (SBO 1024), 64 -> SWIZZLE_64B (SBO 512)
__host__ __device__ constexpr uint64_t cf_desc_hi(int rowb) {
    return ((uint64_t)1 << 16) | ((uint64_t)((rowb * 8) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(rowb == 128 ? 2 : 4) << 61);
}
// byte offset of 16-byte chunk c of position/row p (the swizzle XOR uses absolute address bits; every tile base is 1024-aligned)
template <int ROWB> __device__ __forceinline__ uint32_t cf_swz(int p, int c) {
    return ROWB == 128 ? (uint32_t)p * 128u + (uint32_t)((c ^ (p & 7)) << 4) : (uint32_t)p * 64u + (uint32_t)((c ^ ((p >> 1) & 3)) << 4);
}
// first staged padded row of the tile starting at padded-linear position q0 (floor division, q0 - halo may be negative)
__device__ __forceinline__ int cf_row_lo(int q0, int halo, int Wp) { return (q0 - halo + Wp) / Wp - 1 - halo; }

__device__ __forceinline__ uint4 cf_ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void cf_unpack(const uint4& r, float v[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ uint4 cf_pack(const float v[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void cf_bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cf_bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cf_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cf_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// One 128-position block [q0, q0+128) of image b: lane `lane` owns the interior part of padded row q0/Wp + lane.
// Returns the number of live positions of that row segment (0: none); m0 = index of its first position within the block,
// pix = (b*H + r)*W + c of its first pixel.
__device__ __forceinline__ int cf_segment(const CfGeom& g, int b, int q0, int lane, int& m0, size_t& pix) {
    const int r = (int)__umulhi((unsigned)q0, g.wp_magic) + lane;
    const int rs = r * g.Wp + g.halo;                   // first interior position of the row
    const int lo = max(q0, rs), hi = min(q0 + 128, rs + g.W);
    if (r >= g.H || hi <= lo) return 0;
    m0 = lo - q0;
    pix = ((size_t)b * g.H + r) * g.W + (lo - rs);
    return hi - lo;
}

template <int CIN, int COUT, int NTAPS, int MODE>
__global__ void __launch_bounds__(kCfThreads, 1)
conv_cf_kernel(const __grid_constant__ CUtensorMap tmap_x, const __nv_bfloat16* __restrict__ wp, __nv_bfloat16* __restrict__ y,
               const float* __restrict__ in_scale, const float* __restrict__ in_shift, const __grid_constant__ CfGeom g,
               const __grid_constant__ CfEpi ep) {
    constexpr int ROWB = CIN == 32 ? 64 : 128;                        // bytes per staged position per plane (one swizzle row)
    constexpr int RU = ROWB / 16;                                     // descriptor units (16 B) per row
    constexpr int KC = (CIN + 63) / 64;                               // channel planes per position
    constexpr int KPP = CIN >= 64 ? 4 : CIN / 16;                     // K=16 steps per plane
    constexpr int CH = CIN / 8;                                       // 16-byte channel chunks per position
    constexpr uint32_t W_BYTES = (uint32_t)NTAPS * KC * COUT * ROWB;  // [tap][plane][cout row]
    constexpr int NC = COUT / 2;                                      // channels per epilogue thread
    constexpr uint32_t OUT_BYTES = 128u * COUT * 2;                   // one staged 128-position block of bf16 rows
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // barriers: [0,3) landed (TMA bytes), [3,6) ready (transformed), [6,9) empty, [9,17) tmem_full, [17,25) tmem_empty,
    //           [25,27) out_full, [27,29) out_empty, [29,31) opnd_full
    __shared__ __align__(8) uint64_t bars[31];
    __shared__ uint32_t tmem_slot;
    __shared__ bool is_last;
    __shared__ __align__(16) float cst[MODE == kCfBnRed ? 4 * COUT : COUT];   // statistics: K; bn-backward: mean, invstd, scale, shift
    __shared__ float red[MODE == kCfPlain ? 1 : 8 * COUT];                    // [8 warps][2*NC]
    uint64_t* bar_landed = bars, *bar_ready = bars + 3, *bar_empty = bars + 6, *bar_tfull = bars + 9, *bar_tempty = bars + 17;
    uint64_t* bar_ofull = bars + 25, *bar_oempty = bars + 27, *bar_pfull = bars + 29;

    const int S = g.S, MM = g.MM, NACC = 2 * MM;
    const uint32_t stage_bytes = (uint32_t)KC * g.P * ROWB;           // [plane][position row]
    // every operand tile starts 1024-byte aligned (>= one swizzle period); P % 16 == 0 and Cout % 16 == 0 keep it so
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t w_s = smem_u32(smem);
    const uint32_t a_s = w_s + W_BYTES;
    const uint32_t o_s = a_s + (uint32_t)S * stage_bytes;             // output staging [2][128][COUT]
    const uint32_t p_s = o_s + 2 * OUT_BYTES;                         // operand staging [n_ops][2][128][COUT] (when g.ops_staged)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool xform = in_scale != nullptr;
    const bool has_add = ep.add != nullptr;
    const bool has_out = MODE == kCfBnRed && ep.bn_out != nullptr;
    const int n_ops = (has_add ? 1 : 0) + (MODE == kCfBnRed ? 1 : 0) + (has_out ? 1 : 0);
    const bool staged = g.ops_staged != 0 && n_ops > 0;
    // operand slots in the staging area: z, out, add (only the present ones, in this order)
    const int slot_z = 0, slot_out = (MODE == kCfBnRed ? 1 : 0), slot_add = slot_out + (has_out ? 1 : 0);
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < NACC * COUT) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kCfMaxStages; ++s) {
            mbar_init(smem_u32(bar_landed + s), 1); mbar_init(smem_u32(bar_ready + s), kCfXfThreads / 32); mbar_init(smem_u32(bar_empty + s), 2);
        }
        for (int a = 0; a < 8; ++a) { mbar_init(smem_u32(bar_tfull + a), 1); mbar_init(smem_u32(bar_tempty + a), kCfEpiThreads / 32); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(bar_ofull + a), kCfEpiThreads / 32); mbar_init(smem_u32(bar_oempty + a), 1); mbar_init(smem_u32(bar_pfull + a), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), tmem_cols);
    // weights: global (tap, co, ci) at co*w_row_stride + tap*w_tap_stride + ci -> smem [tap][plane][co] rows of ROWB bytes, swizzled (K-major B operand)
    for (int i = threadIdx.x; i < NTAPS * COUT * CH; i += kCfThreads) {
        const int kc = i % CH, row = i / CH, co = row % COUT, tap = row / COUT;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(wp + ((size_t)co * g.w_row_stride + (size_t)tap * g.w_tap_stride + kc * 8)));
        const int plane = kc >> 3, c = kc & 7;
        *reinterpret_cast<uint4*>(smem + (size_t)(tap * KC + plane) * COUT * ROWB + cf_swz<ROWB>(co, c)) = v;
    }
    // everything above touched only this CTA's shared memory / TMEM and the weight shadow (written by the PREVIOUS step's optimiser):
    // it may overlap the tail of the preceding kernel.  From here on the kernel reads what its predecessors produced.
    pdl_wait();
    pdl_trigger();
    if (MODE == kCfStats) {
        for (int c = threadIdx.x; c < COUT; c += kCfThreads) cst[c] = ep.running_mean ? ep.running_mean[c] : 0.f;
    } else if (MODE == kCfBnRed) {
        for (int c = threadIdx.x; c < COUT; c += kCfThreads) {
            cst[c] = ep.bn_mean[c]; cst[COUT + c] = ep.bn_invstd[c]; cst[2 * COUT + c] = ep.bn_scale[c]; cst[3 * COUT + c] = ep.bn_shift[c];
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: NR padded rows x Wp pixels x one channel plane per box, OOB zero fill = padding =================
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)KC * g.NR * g.Wp * ROWB;
            int si = 0, use = 0, ti = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++ti) {
                CF_TRACE(0, ti, 0);
                if (use > 0) mbar_wait(smem_u32(bar_empty + si), (use - 1) & 1);         // MMAs that read this stage retired
                CF_TRACE(0, ti, 1);
                const int b = tile / g.tiles_per_img, t = tile - b * g.tiles_per_img;
                const int r_lo = cf_row_lo(t * g.MT, g.halo, g.Wp);
                const uint32_t full = smem_u32(bar_landed + si);
                mbar_expect_tx(full, tx_bytes);
                if (g.xsplit == 1) {
#pragma unroll
                    for (int pl = 0; pl < KC; ++pl)
                        tma_load_4d(a_s + si * stage_bytes + (uint32_t)(pl * g.P) * ROWB, &tmap_x, full, pl * 64, -g.halo, r_lo, b);
                } else {                                                        // wide images: one box per row and column part
                    for (int pl = 0; pl < KC; ++pl)
                        for (int rr = 0; rr < g.NR; ++rr)
                            for (int xs = 0; xs < g.xsplit; ++xs)
                                tma_load_4d(a_s + si * stage_bytes + (uint32_t)(pl * g.P + rr * g.Wp + xs * g.box_w) * ROWB, &tmap_x, full,
                                            pl * 64, -g.halo + xs * g.box_w, r_lo + rr, b);
                }
                if (++si == S) { si = 0; ++use; }
            }
        }
    } else if (warp == 1 || warp == 15) {
        // ================= MMA issuers: warp 1 issues the even 128-row blocks of every tile, warp 15 the odd ones.  One warp cannot
        // keep the tensor pipe fed: it sits in back-pressure while its MMAs execute (32 cycles each: the 4 KB A operand at 128 B/clk)
        // and only then does the bookkeeping of the next block (~450 cycles of barrier polls and descriptor arithmetic per block).
        // The whole warp runs the (warp-uniform) loop, one elected lane issues. =================
        const int who = warp == 1 ? 0 : 1;
        constexpr uint32_t idesc = cf_idesc(COUT);
        uint64_t* bar_in = xform ? bar_ready : bar_landed;
        constexpr uint64_t desc_hi = cf_desc_hi(ROWB);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t leader = elect_one();
        const uint32_t w_lo = w_s >> 4;
        const int plane_a = g.P * RU;
        const bool no_mma = (g.dbg & 4) != 0;
        int toff[NTAPS];
#pragma unroll
        for (int t = 0; t < NTAPS; ++t) toff[t] = g.tap_off[t] * RU;
        int si = 0, ti = 0;
        uint32_t in_phase = 0;
        // accumulator of block mm of the tile: (tile parity)*MM + mm; its phase flips every second tile
        int tpar = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++ti) {
            const int b_ = tile / g.tiles_per_img;
            const int q0 = (tile - b_ * g.tiles_per_img) * g.MT;
            const int pbase = q0 - cf_row_lo(q0, g.halo, g.Wp) * g.Wp;  // staged index of output position q0
            // descriptors differ only in their 14-bit start-address field (address >> 4): a K=16 step of 32 B = 2 units
            const int a_lo0 = (int)((a_s + si * stage_bytes) >> 4) + pbase * RU;
            if (lane == 0 && who == 0) CF_TRACE(1, ti, 0);
            mbar_wait(smem_u32(bar_in + si), in_phase);                // tile staged (and transformed)
            if (lane == 0 && who == 0) CF_TRACE(1, ti, 1);
            tc_fence_after();
            for (int mm = who; mm < MM; mm += 2) {
                const int acc = tpar * MM + mm;
                mbar_wait(smem_u32(bar_tempty + acc), acc_phase ^ 1);  // epilogue drained this accumulator
                if (lane == 0 && mm < 2) CF_TRACE(1, ti, 2 + 2 * mm);
                tc_fence_after();
                const uint32_t d_tmem = tmem_u + acc * COUT;
                const int a_lo1 = a_lo0 + mm * 128 * RU;
                if (!no_mma) {
#pragma unroll
                    for (int tp = 0; tp < NTAPS; ++tp) {
#pragma unroll
                        for (int pl = 0; pl < KC; ++pl) {
#pragma unroll
                            for (int kk = 0; kk < KPP; ++kk) {
                                const uint32_t a = (uint32_t)(a_lo1 + toff[tp] + pl * plane_a + kk * 2);
                                const uint32_t b = w_lo + (uint32_t)((tp * KC + pl) * COUT * RU + kk * 2);
                                umma_bf16_elect(leader, d_tmem, desc_hi | (uint64_t)a, desc_hi | (uint64_t)b, idesc, (tp | pl | kk) ? 1u : 0u);
                            }
                        }
                    }
                }
                umma_commit_elect(leader, smem_u32(bar_tfull + acc));  // accumulator complete -> epilogue
                if (lane == 0 && mm < 2) CF_TRACE(1, ti, 3 + 2 * mm);
            }
            umma_commit_elect(leader, smem_u32(bar_empty + si));        // stage free once both issuers' MMAs on it retire
            if (++si == S) { si = 0; in_phase ^= 1; }
            if (++tpar == 2) { tpar = 0; acc_phase ^= 1; }
        }
    } else if (warp < 6) {
        // ================= transform warps: previous layer's BN(+ReLU) applied in place on the landed tile =================
        if (xform) {
            const int ht = threadIdx.x - 64;                            // 0..127
            const int ch = ht % CH, c_first = ht / CH;                  // this thread's channel chunk is fixed (CH divides 128)
            constexpr int PSTEP = kCfXfThreads / CH;
            float sc[8], sh[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { sc[k] = in_scale[ch * 8 + k]; sh[k] = in_shift[ch * 8 + k]; }
            const bool relu = g.in_relu != 0;
            int si = 0, ti = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++ti) {
                if (ht == 0) CF_TRACE(2, ti, 0);
                mbar_wait(smem_u32(bar_landed + si), phase);
                if (ht == 0) CF_TRACE(2, ti, 1);
                const int t = tile % g.tiles_per_img;
                const int r_lo = cf_row_lo(t * g.MT, g.halo, g.Wp);
                uint8_t* base = smem + W_BYTES + (size_t)si * stage_bytes + (size_t)((ch >> 3) * g.P) * ROWB;
                for (int rr = 0; rr < g.NR; ++rr) {
                    const int r = r_lo + rr;
                    if (r < 0 || r >= g.H) continue;                    // padding rows stay exactly zero
                    const int p0 = rr * g.Wp + g.halo;                  // interior columns only: padding columns stay zero
                    for (int cb = c_first; cb < g.W; cb += 4 * PSTEP) {            // 4 chunks in flight per thread
                        uint4* ptr[4];
                        uint4 raw[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int c = cb + u * PSTEP;
                            ptr[u] = reinterpret_cast<uint4*>(base + cf_swz<ROWB>(p0 + c, ch & 7));
                            if (c < g.W) raw[u] = *ptr[u];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float v[8];
                            cf_unpack(raw[u], v);
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                v[k] = fmaf(v[k], sc[k], sh[k]);
                                if (relu) v[k] = fmaxf(v[k], 0.f);
                            }
                            if (cb + u * PSTEP < g.W) *ptr[u] = cf_pack(v);
                        }
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(bar_ready + si));
                if (ht == 0) CF_TRACE(2, ti, 2);
                if (++si == S) { si = 0; phase ^= 1; }
            }
        }
    } else if (warp < 14) {
        // ================= epilogue: TMEM -> (+add, mask, reductions) -> bf16 rows in the staging buffer =================
        const int ew = warp - 6, grp = ew >> 2;                // group 0: channels [0, NC), group 1: [NC, COUT)
        const int q4 = warp & 3;                               // TMEM lane quarter this warp may access
        const int m = q4 * 32 + lane;
        const int ch0 = grp * NC;
        constexpr int NS = MODE == kCfPlain ? 1 : NC;
        float s1[NS], s2[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
        const bool bn_relu = MODE == kCfBnRed && ep.bn_relu != 0;
        const uint32_t row_off = (uint32_t)m * COUT * 2 + (uint32_t)ch0 * 2;      // this thread's bytes within a staged block
        int ti = 0, kb = 0, tpar = 0;
        uint32_t acc_phase = 0;
        const bool tracer = threadIdx.x == 192;
        const bool no_epi = (g.dbg & 1) != 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++ti) {
            const int b = tile / g.tiles_per_img, t = tile - b * g.tiles_per_img;
            for (int mm = 0; mm < MM; ++mm, ++kb) {
                if (tracer && mm < 2) CF_TRACE(3, ti, 4 * mm);
                const int ob = kb & 1;
                const uint32_t ph = (uint32_t)(kb >> 1) & 1;
                const int acc = tpar * MM + mm;
                const int q = t * g.MT + mm * 128 + m;
                const int r = (int)__umulhi((unsigned)q, g.wp_magic), c = q - r * g.Wp - g.halo;
                const bool live = q < g.Q && c >= 0 && c < g.W && !no_epi;
                uint4 av[NC / 8], zv[NC / 8], ov[NC / 8];
                if (!staged && n_ops > 0 && live) {            // direct (strided) operand loads: in flight while the MMAs finish
                    const size_t off = (((size_t)b * g.H + r) * g.W + c) * COUT + ch0;
                    if (has_add) {
#pragma unroll
                        for (int i = 0; i < NC / 8; ++i) av[i] = cf_ldg16(ep.add + off + i * 8);
                    }
                    if (MODE == kCfBnRed) {
#pragma unroll
                        for (int i = 0; i < NC / 8; ++i) zv[i] = cf_ldg16(ep.bn_z + off + i * 8);
                        if (has_out) {
#pragma unroll
                            for (int i = 0; i < NC / 8; ++i) ov[i] = cf_ldg16(ep.bn_out + off + i * 8);
                        }
                    }
                }
                mbar_wait(smem_u32(bar_tfull + acc), acc_phase);
                if (tracer && mm < 2) CF_TRACE(3, ti, 4 * mm + 1);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * COUT + ch0;
                uint32_t rr[NC];
#pragma unroll
                for (int c0 = 0; c0 < NC; c0 += 16) tmem_ld16(t_row + c0, rr + c0);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(bar_tempty + acc));      // the accumulator is in registers: hand TMEM back
                if (staged) {                                  // operand rows brought in by the copy warp
                    mbar_wait(smem_u32(bar_pfull + ob), ph);
                    const uint8_t* pb = smem + (p_s - w_s) + (size_t)ob * OUT_BYTES + row_off;
                    if (has_add) {
#pragma unroll
                        for (int i = 0; i < NC / 8; ++i) av[i] = *reinterpret_cast<const uint4*>(pb + (size_t)slot_add * 2 * OUT_BYTES + i * 16);
                    }
                    if (MODE == kCfBnRed) {
#pragma unroll
                        for (int i = 0; i < NC / 8; ++i) zv[i] = *reinterpret_cast<const uint4*>(pb + (size_t)slot_z * 2 * OUT_BYTES + i * 16);
                        if (has_out) {
#pragma unroll
                            for (int i = 0; i < NC / 8; ++i) ov[i] = *reinterpret_cast<const uint4*>(pb + (size_t)slot_out * 2 * OUT_BYTES + i * 16);
                        }
                    }
                }
                if (tracer && mm < 2) CF_TRACE(3, ti, 4 * mm + 2);
                uint4 pk[NC / 8];
                const float livef = live ? 1.f : 0.f;          // everything below is branch-free (per-element branches diverge)
#pragma unroll
                for (int i = 0; i < NC / 8; ++i) {
                    float v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(rr[i * 8 + k]);
                    if (has_add) {
                        float a8[8];
                        cf_unpack(av[i], a8);
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] += a8[k];
                    }
                    if (MODE == kCfStats) {                    // statistics of the fp32 values (rounding is zero-mean)
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float d = (v[k] - cst[ch0 + i * 8 + k]) * livef;
                            s1[i * 8 + k] += d;
                            s2[i * 8 + k] = fmaf(d, d, s2[i * 8 + k]);
                        }
                    } else if (MODE == kCfBnRed) {
                        float z8[8], o8[8];
                        cf_unpack(zv[i], z8);
                        if (has_out) cf_unpack(ov[i], o8);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int cc = ch0 + i * 8 + k;
                            const float zk = live ? z8[k] : 0.f;              // rows of dead positions hold garbage
                            const float vk = live ? v[k] : 0.f;
                            float msrc = 1.f;
                            if (has_out) msrc = live ? o8[k] : 0.f;
                            else if (bn_relu) msrc = fmaf(zk, cst[2 * COUT + cc], cst[3 * COUT + cc]);
                            const float gk = msrc > 0.f ? vk : 0.f;
                            const float xh = (zk - cst[cc]) * cst[COUT + cc];
                            s1[i * 8 + k] += gk;
                            s2[i * 8 + k] = fmaf(gk, xh, s2[i * 8 + k]);
                            v[k] = gk;
                        }
                    }
                    pk[i] = cf_pack(v);
                }
                mbar_wait(smem_u32(bar_oempty + ob), ph ^ 1);  // the bulk stores of block kb-2 have read this staging buffer
                if (live) {
                    uint8_t* ob_ptr = smem + (o_s - w_s) + (size_t)ob * OUT_BYTES + row_off;
#pragma unroll
                    for (int i = 0; i < NC / 8; ++i) *reinterpret_cast<uint4*>(ob_ptr + i * 16) = pk[i];
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(bar_ofull + ob));
                if (tracer && mm < 2) CF_TRACE(3, ti, 4 * mm + 3);
            }
            if (++tpar == 2) { tpar = 0; acc_phase ^= 1; }
        }
        if (MODE != kCfPlain) {
            // per-channel totals: recursive-halving butterfly (NS-1 shuffles per array instead of 5*NS: after the step with offset o
            // a lane keeps the half of its values selected by its lane bit o) -> 8 warp partials in smem -> one atomicAdd per
            // channel per CTA.  Lane l ends up with the warp total of value index bitrev-free: idx = sum of kept halves.
            {
                int idx = 0;                                   // value index this lane ends up owning
                int n = NS;
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    if (n > 1) {
                        n >>= 1;
                        const bool up = (lane & o) != 0;       // upper lanes keep the upper half of the remaining values
#pragma unroll
                        for (int i = 0; i < NS / 2; ++i) {
                            if (i < n) {
                                const float keep1 = up ? s1[n + i] : s1[i], send1 = up ? s1[i] : s1[n + i];
                                const float keep2 = up ? s2[n + i] : s2[i], send2 = up ? s2[i] : s2[n + i];
                                s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, o);
                                s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, o);
                            }
                        }
                        if (up) idx += n;
                    } else {                                   // one value left: plain butterfly over the remaining lane bits
                        s1[0] += __shfl_xor_sync(0xffffffffu, s1[0], o);
                        s2[0] += __shfl_xor_sync(0xffffffffu, s2[0], o);
                    }
                }
                // lanes that differ only in the bits consumed by the plain butterfly hold the same total: one of them writes
                constexpr int LANES_PER_VAL = 32 / (NS < 32 ? NS : 32);
                if ((lane & (LANES_PER_VAL - 1)) == 0) { red[ew * 2 * NC + idx] = s1[0]; red[ew * 2 * NC + NC + idx] = s2[0]; }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int et = threadIdx.x - 192;                  // 0..255 within the epilogue group
            for (int i = et; i < 2 * COUT; i += kCfEpiThreads) {
                const int which = i / COUT, chn = i - which * COUT, gg = chn / NC, cl = chn - gg * NC;
                const float* rp = red + (gg * 4) * 2 * NC + which * NC + cl;
                atomicAdd(ep.accum + i, rp[0] + rp[2 * NC] + rp[4 * NC] + rp[6 * NC]);
            }
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et == 0) is_last = (atomicAdd(ep.ticket, 1u) == gridDim.x - 1);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (is_last) {
                __threadfence();
                for (int c = et; c < COUT; c += kCfEpiThreads) {
                    const float Ssum = __ldcg(ep.accum + c), Qsum = __ldcg(ep.accum + COUT + c);
                    if (MODE == kCfStats) {
                        const float md = Ssum / ep.count;
                        const float m2 = fmaxf(Qsum - Ssum * md, 0.f);
                        const float mean = cst[c] + md;
                        const float invstd = rsqrtf(m2 / ep.count + ep.eps);
                        ep.mean_out[c] = mean;
                        ep.invstd_out[c] = invstd;
                        const float scl = ep.gamma[c] * invstd;
                        ep.scale_out[c] = scl;
                        ep.shift_out[c] = ep.beta[c] - mean * scl;
                        if (ep.running_mean) {
                            ep.running_mean[c] = (1.f - ep.momentum) * ep.running_mean[c] + ep.momentum * mean;
                            ep.running_var[c] = (1.f - ep.momentum) * ep.running_var[c] + ep.momentum * (m2 / fmaxf(ep.count - 1.f, 1.f));
                        }
                    } else {
                        ep.sums_out[c] = Ssum;
                        ep.sums_out[COUT + c] = Qsum;
                    }
                    ep.accum[c] = 0.f;
                    ep.accum[COUT + c] = 0.f;
                }
                if (et == 0) *ep.ticket = 0u;
            }
        }
    } else {
        // ================= copy warp: staged output rows -> global (bulk stores), epilogue operands -> staging (bulk loads) =================
        // lane l handles the row segment of padded row q0/Wp + l of each 128-position block (<= 32 rows per block: Wp >= 8)
        const __nv_bfloat16* op_ptr[3];
        op_ptr[slot_z] = ep.bn_z;
        if (has_out) op_ptr[slot_out] = ep.bn_out;
        if (has_add) op_ptr[slot_add] = ep.add;
        // (tile2, mm2): the block two ahead of the one being stored -- its operands are fetched into the stage just released
        int tile2 = blockIdx.x, mm2 = 0, kb2 = 0;
        auto load_ops = [&](int tl, int mmx, int kbx) {
            const int b = tl / g.tiles_per_img, t = tl - b * g.tiles_per_img;
            int m0 = 0; size_t pix = 0;
            const int n = cf_segment(g, b, t * g.MT + mmx * 128, lane, m0, pix);
            const uint32_t bytes = (uint32_t)n * COUT * 2;
            const uint32_t total = __reduce_add_sync(0xffffffffu, bytes) * (uint32_t)n_ops;
            const int ob = kbx & 1;
            const uint32_t bar = smem_u32(bar_pfull + ob);
            if (lane == 0) mbar_expect_tx(bar, total);
            __syncwarp();
            if (n > 0) {
                for (int o = 0; o < n_ops; ++o)
                    cf_bulk_load(p_s + (uint32_t)(o * 2 + ob) * OUT_BYTES + (uint32_t)m0 * COUT * 2, op_ptr[o] + pix * COUT, bytes, bar);
            }
        };
        auto advance2 = [&]() { ++kb2; if (++mm2 == MM) { mm2 = 0; tile2 += gridDim.x; } };
        if (staged) {
            for (int i = 0; i < 2 && tile2 < g.n_tiles; ++i) { load_ops(tile2, mm2, kb2); advance2(); }
        }
        int kb = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
            const int b = tile / g.tiles_per_img, t = tile - b * g.tiles_per_img;
            for (int mm = 0; mm < MM; ++mm, ++kb) {
                const int ob = kb & 1;
                const uint32_t ph = (uint32_t)(kb >> 1) & 1;
                int m0 = 0; size_t pix = 0;
                const int n = cf_segment(g, b, t * g.MT + mm * 128, lane, m0, pix);
                mbar_wait(smem_u32(bar_ofull + ob), ph);       // all 8 epilogue warps wrote (and fenced) their rows of block kb
                if (n > 0) cf_bulk_store(y + pix * COUT, o_s + (uint32_t)ob * OUT_BYTES + (uint32_t)m0 * COUT * 2, (uint32_t)n * COUT * 2);
                cf_bulk_commit();
                if (staged && tile2 < g.n_tiles) { load_ops(tile2, mm2, kb2); advance2(); }    // block kb's operands are consumed
                cf_bulk_wait_read0();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(bar_oempty + ob));
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // global writes of this CTA complete before it exits
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

struct CfPlan { CfGeom g; size_t smem; int grid; };

static bool cf_shape_ok(int Cin, int Cout, int n_taps, int mode);

static int cf_plan(int B, int H, int W, int Cin, int Cout, int n_taps, const int* dy, const int* dx, int n_ops, CfPlan* pl) {
    if (B <= 0 || H <= 0 || W <= 0 || n_taps < 1 || n_taps > kCfMaxTaps) return RSS_ERR_SHAPE;
    CfGeom& g = pl->g;
    g.B = B; g.H = H; g.W = W; g.in_relu = 0; g.trace = nullptr; g.ops_staged = 0;
    int halo = 0;
    for (int t = 0; t < n_taps; ++t) {
        const int a = dy[t] < 0 ? -dy[t] : dy[t], b = dx[t] < 0 ? -dx[t] : dx[t];
        if (a > halo) halo = a;
        if (b > halo) halo = b;
    }
    if (halo > 1) return RSS_ERR_SHAPE;
    g.halo = halo; g.Wp = W + 2 * halo;
    if (g.Wp > 256) g.Wp = (g.Wp + 15) & ~15;      // two TMA boxes per staged row (below): both must start on a swizzle-atom boundary
    g.Q = H * g.Wp;
    g.wp_magic = (unsigned int)((0x100000000ull + (unsigned)g.Wp - 1) / (unsigned)g.Wp);
    g.dbg = 0;
    if (g.Wp < 8) return RSS_ERR_SHAPE;                                       // <= 32 row segments per 128-position block
    g.xsplit = 1; g.box_w = g.Wp;
    if (g.Wp > 256) {                                                         // TMA box dimension limit: split every row in two boxes
        if (g.Wp > 512) return RSS_ERR_SHAPE;                                 // (the pitch was rounded up to 16 positions: the extra
        g.xsplit = 2; g.box_w = g.Wp / 2;                                     //  columns are dead positions like the padding ones)
    }
    for (int t = 0; t < kCfMaxTaps; ++t) g.tap_off[t] = t < n_taps ? dy[t] * g.Wp + dx[t] : 0;
    const int KC = (Cin + 63) / 64, rowb = Cin == 32 ? 64 : 128;
    const size_t w_bytes = (size_t)n_taps * KC * Cout * rowb;
    const size_t out_bytes = (size_t)128 * Cout * 2;
    const size_t budget = 227 * 1024 - 8 * 1024 - 1024;                       // static shared memory (<= 6.5 KB) + manual alignment
    // two 128-row blocks per tile halve the halo over-fetch; needs 4 accumulators in TMEM and a tile count that still fills the GPU.
    // Preference: staged epilogue operands > two blocks per tile > three ring stages.
    for (int staged = n_ops > 0 ? 1 : 0; staged >= 0; --staged) {
        for (int mm = 4; mm >= 1; mm >>= 1) {
            if (2 * mm * Cout > 512) continue;                                // 2*mm accumulators of Cout TMEM columns
            g.MM = mm; g.MT = 128 * mm;
            g.tiles_per_img = (g.Q + g.MT - 1) / g.MT;
            if (mm > 1 && (int64_t)B * g.tiles_per_img < 2 * num_sms()) continue;
            const int L = g.MT + 2 * halo;                                    // padded-linear span a tile reads within its own rows
            g.NR = (L + g.Wp - 2) / g.Wp + 1 + 2 * halo;                      // rows that span can touch, + halo rows above and below
            if (g.NR > 256) continue;
            g.P = (g.NR * g.Wp + 15) & ~15;                                   // planes/stages stay 1024-aligned for both row sizes
            const size_t stage = (size_t)KC * g.P * rowb;
            for (int s = kCfMaxStages; s >= 2; --s) {
                const size_t need = w_bytes + (size_t)s * stage + 2 * out_bytes + (staged ? (size_t)n_ops * 2 * out_bytes : 0);
                if (need <= budget) {
                    g.S = s; pl->smem = need + 1024;
                    g.ops_staged = staged;
                    g.n_tiles = B * g.tiles_per_img;
                    pl->grid = g.n_tiles < num_sms() ? g.n_tiles : num_sms();
                    return RSS_OK;
                }
            }
        }
    }
    return RSS_ERR_SHAPE;
}

typedef CUresult (*CfEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CfEncodeTiledFn cf_encode_tiled() {
    static CfEncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (CfEncodeTiledFn)p;
    }
    return fn;
}

// ---- the instantiated geometries: (Cin, Cout, taps) x epilogue mode ----
//   3x3: 32->32 (branch 0), 64->64 (branch 1, layer1 conv2): all three epilogues
//   1x1: 64->32, 128->32, 128->64, 64->64, 32->128: plain (+ statistics when Cout <= 64)
#define CF_FOR_EACH_SHAPE(X) \
    X(32, 32, 9, 0) X(32, 32, 9, 1) X(32, 32, 9, 2) X(64, 64, 9, 0) X(64, 64, 9, 1) X(64, 64, 9, 2) \
    X(64, 32, 1, 0) X(64, 32, 1, 1) X(128, 32, 1, 0) X(128, 32, 1, 1) X(128, 64, 1, 0) X(128, 64, 1, 1) \
    X(64, 64, 1, 0) X(64, 64, 1, 1) X(32, 128, 1, 0)

static bool cf_shape_ok(int Cin, int Cout, int n_taps, int mode) {
#define X(CI, CO, NT, MD) if (Cin == CI && Cout == CO && n_taps == NT && mode == MD) return true;
    CF_FOR_EACH_SHAPE(X)
#undef X
    return false;
}

template <int CIN, int COUT, int NTAPS, int MODE>
static cudaError_t cf_launch(const CUtensorMap& tm, const void* w_packed, void* y, const float* in_scale, const float* in_shift,
                             const CfPlan& pl, const CfEpi& ep, cudaStream_t stream) {
    static bool attr[16] = {};                                                // per device: the opt-in is a per-device function attribute
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) return cudaErrorInvalidDevice;
    if (!attr[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_cf_kernel<CIN, COUT, NTAPS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(227 * 1024 - 8 * 1024));
        if (e != cudaSuccess) return e;
        attr[dev] = true;
    }
    launch_k(conv_cf_kernel<CIN, COUT, NTAPS, MODE>, pl.grid, kCfThreads, pl.smem, stream,
             tm, (const __nv_bfloat16*)w_packed, (__nv_bfloat16*)y, in_scale, in_shift, pl.g, ep);
    return cudaSuccess;
}

}  // namespace rss

using namespace rss;

// 1 when rss_conv_cf accepts the geometry.  mode: RSS_CF_PLAIN / RSS_CF_STATS / RSS_CF_BNRED
extern "C" int rss_conv_cf_supported(int B, int H, int W, int Cin, int Cout, int ksize, int mode) {
    if (ksize != 1 && ksize != 3) return 0;
    if (!cf_shape_ok(Cin, Cout, ksize * ksize, mode)) return 0;
    int dy[9], dx[9], n = 0;
    for (int a = 0; a < ksize; ++a) for (int b = 0; b < ksize; ++b) { dy[n] = a - ksize / 2; dx[n] = b - ksize / 2; ++n; }
    CfPlan pl;
    return cf_plan(B, H, W, Cin, Cout, n, dy, dx, 0, &pl) == RSS_OK;
}

// y = E(conv(T(x)) + add); weight operand bf16, element (tap, co, ci) at co*w_row_stride + tap*w_tap_stride + ci (both strides
// 0: the [tap][Cout][Cin] pack of rss_conv_pack_weights, forward or transposed), taps (dy,dx) within [-1,1].  in_scale/in_shift (fp32 [Cin], NULL = identity) and in_relu describe T; `e` (may be NULL) the epilogue.
extern "C" int rss_conv_cf(const void* x, const void* w_packed, void* y, int B, int H, int W, int Cin, int Cout,
                           int n_taps, const int* taps_dy, const int* taps_dx,
                           int w_row_stride, int w_tap_stride,
                           const float* in_scale, const float* in_shift, int in_relu,
                           const RssConvCfEpilogue* e, cudaStream_t stream) {
    const int mode = e ? e->mode : RSS_CF_PLAIN;
    if (mode != RSS_CF_PLAIN && mode != RSS_CF_STATS && mode != RSS_CF_BNRED) return RSS_ERR_SHAPE;
    if (!cf_shape_ok(Cin, Cout, n_taps, mode)) return RSS_ERR_SHAPE;
    CfPlan pl;
    const int n_ops = e ? (e->add ? 1 : 0) + (mode == RSS_CF_BNRED ? 1 + (e->bn_out ? 1 : 0) : 0) : 0;
    int rc = cf_plan(B, H, W, Cin, Cout, n_taps, taps_dy, taps_dx, n_ops, &pl);
    if (rc != RSS_OK) return rc;
    if ((in_scale == nullptr) != (in_shift == nullptr)) return RSS_ERR_SHAPE;
    if (((uintptr_t)x & 15) || ((uintptr_t)y & 15) || ((uintptr_t)w_packed & 15)) return RSS_ERR_SHAPE;
    pl.g.in_relu = in_relu;
    if (w_row_stride <= 0 && w_tap_stride <= 0) { w_row_stride = Cin; w_tap_stride = Cout * Cin; }     // packed [tap][Cout][Cin]
    if ((w_row_stride & 7) || (w_tap_stride & 7)) return RSS_ERR_SHAPE;                                // 16-byte chunks
    pl.g.w_row_stride = w_row_stride; pl.g.w_tap_stride = w_tap_stride;
    {
        const char* dbg = getenv("RSS_CF_DBG");
        pl.g.dbg = dbg ? atoi(dbg) : 0;
        const char* tr = getenv("RSS_CF_TRACE_PTR");
        pl.g.trace = tr ? (long long*)strtoull(tr, nullptr, 16) : nullptr;
    }
    CfEpi ep{};
    if (e) {
        ep.add = (const __nv_bfloat16*)e->add;
        if ((uintptr_t)e->add & 15) return RSS_ERR_SHAPE;
        ep.accum = e->accum; ep.ticket = e->ticket; ep.count = (float)((double)B * H * W);
        if (mode != RSS_CF_PLAIN && (!e->accum || !e->ticket)) return RSS_ERR_SHAPE;
        if (mode == RSS_CF_STATS) {
            if (!e->gamma || !e->beta || !e->mean_out || !e->invstd_out || !e->scale_out || !e->shift_out) return RSS_ERR_SHAPE;
            ep.gamma = e->gamma; ep.beta = e->beta; ep.running_mean = e->running_mean; ep.running_var = e->running_var;
            ep.momentum = e->momentum; ep.eps = e->eps; ep.mean_out = e->mean_out; ep.invstd_out = e->invstd_out;
            ep.scale_out = e->scale_out; ep.shift_out = e->shift_out;
        } else if (mode == RSS_CF_BNRED) {
            if (!e->bn_z || !e->bn_mean || !e->bn_invstd || !e->bn_scale || !e->bn_shift || !e->sums_out) return RSS_ERR_SHAPE;
            if (((uintptr_t)e->bn_z & 15) || ((uintptr_t)e->bn_out & 15)) return RSS_ERR_SHAPE;
            ep.bn_z = (const __nv_bfloat16*)e->bn_z; ep.bn_out = (const __nv_bfloat16*)e->bn_out; ep.bn_mean = e->bn_mean;
            ep.bn_invstd = e->bn_invstd; ep.bn_scale = e->bn_scale; ep.bn_shift = e->bn_shift; ep.bn_relu = e->bn_relu;
            ep.sums_out = e->sums_out;
        }
    }
    CUtensorMap tm;
    {
        CfEncodeTiledFn enc = cf_encode_tiled();
        if (!enc) return RSS_ERR_CUDA;
        const CfGeom& g = pl.g;
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        cuuint32_t box[4] = {(cuuint32_t)(Cin == 32 ? 32 : 64), (cuuint32_t)g.box_w, (cuuint32_t)(g.xsplit == 1 ? g.NR : 1), 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, Cin == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return RSS_ERR_CUDA; }
    }
    cudaError_t err = cudaErrorInvalidValue;
#define X(CI, CO, NT, MD) \
    if (Cin == CI && Cout == CO && n_taps == NT && mode == MD) err = cf_launch<CI, CO, NT, MD>(tm, w_packed, y, in_scale, in_shift, pl, ep, stream); else
    CF_FOR_EACH_SHAPE(X) {}
#undef X
    if (err != cudaSuccess) { g_last_cuda_error = (int)err; (void)cudaGetLastError(); return RSS_ERR_CUDA; }
    return check_launch();
}
