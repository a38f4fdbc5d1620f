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

constexpr int kCfThreads = 448;          // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 input transform, warps 6-13 epilogue
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
    int NR, P;                           // staged padded rows per tile; plane pitch in positions (>= NR*Wp, multiple of 8)
    int S;                               // ring stages (2 or 3)
    int tap_off[kCfMaxTaps];             // dy*Wp + dx (signed)
    int in_relu;
    long long* trace;                    // profiling aid (RSS_CF_TRACE_PTR): CTA 0 records clock64() per role/tile/event, [4 roles][16 tiles][8]
};
#define CF_TRACE(role, i, k) do { if (g.trace && blockIdx.x == 0 && (i) < 16) g.trace[((role) * 16 + (i)) * 8 + (k)] = clock64(); } while (0)

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
__device__ __forceinline__ void cf_epi_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// first staged padded row of the tile starting at padded-linear position q0 (floor division, q0 - halo may be negative)
__device__ __forceinline__ int cf_row_lo(int q0, int halo, int Wp) { return (q0 - halo + Wp) / Wp - 1 - halo; }

__device__ __forceinline__ uint4 cf_ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void cf_unpack(const uint4& r, float v[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}

template <int CIN, int COUT, int NTAPS, int MODE>
__global__ void __launch_bounds__(kCfThreads, 1)
conv_cf_kernel(const __grid_constant__ CUtensorMap tmap_x, const __nv_bfloat16* __restrict__ wp, __nv_bfloat16* __restrict__ y,
               const float* __restrict__ in_scale, const float* __restrict__ in_shift, const __grid_constant__ CfGeom g,
               const __grid_constant__ CfEpi ep) {
    constexpr int KC = (CIN + 63) / 64;                               // 64-channel planes per position
    constexpr int KPP = CIN >= 64 ? 4 : CIN / 16;                     // K=16 steps per plane
    constexpr int CH = CIN / 8;                                       // 16-byte channel chunks per position
    constexpr uint32_t W_BYTES = (uint32_t)NTAPS * KC * COUT * 128;   // [tap][plane][cout row of 128 B]
    constexpr int NC = COUT / 2;                                      // channels per epilogue thread
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // barriers: [0,3) landed (TMA bytes), [3,6) ready (transformed), [6,9) empty, [9,13) tmem_full, [13,17) tmem_empty
    __shared__ __align__(8) uint64_t bars[17];
    __shared__ uint32_t tmem_slot;
    __shared__ bool is_last;
    __shared__ __align__(16) float cst[MODE == kCfBnRed ? 4 * COUT : COUT];   // statistics: K; bn-backward: mean, invstd, scale, shift
    __shared__ float red[MODE == kCfPlain ? 1 : 8 * COUT];                    // [8 warps][2*NC]
    uint64_t* bar_landed = bars, *bar_ready = bars + 3, *bar_empty = bars + 6, *bar_tfull = bars + 9, *bar_tempty = bars + 13;

    const int S = g.S, MM = g.MM, NACC = 2 * MM;
    const uint32_t stage_bytes = (uint32_t)KC * g.P * 128;            // [plane][position row of 128 B]
    // every operand tile starts 1024-byte aligned (one swizzle period); P % 8 == 0 and Cout % 8 == 0 keep it so
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t w_s = smem_u32(smem);
    const uint32_t a_s = w_s + W_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool xform = in_scale != nullptr;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < NACC * COUT) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kCfMaxStages; ++s) {
            mbar_init(smem_u32(bar_landed + s), 1); mbar_init(smem_u32(bar_ready + s), kCfXfThreads / 32); mbar_init(smem_u32(bar_empty + s), 1);
        }
        for (int a = 0; a < 4; ++a) { mbar_init(smem_u32(bar_tfull + a), 1); mbar_init(smem_u32(bar_tempty + a), kCfEpiThreads / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), tmem_cols);
    // weights: packed global [tap][co][ci] -> smem [tap][plane][co][64 ci] rows of 128 B, 128B-swizzled (K-major B operand)
    for (int i = threadIdx.x; i < NTAPS * COUT * CH; i += kCfThreads) {
        const int kc = i % CH, row = i / CH, co = row % COUT, tap = row / COUT;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(wp + ((size_t)row * CIN + kc * 8)));
        const int plane = kc >> 3, c = kc & 7;
        *reinterpret_cast<uint4*>(smem + ((size_t)(tap * KC + plane) * COUT + co) * 128 + ((c ^ (co & 7)) << 4)) = v;
    }
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
        // ================= TMA producer: NR padded rows x Wp pixels x 64 channels per plane, OOB zero fill = padding =================
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)KC * g.NR * g.Wp * 128;
            int si = 0, use = 0, ti = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++ti) {
                CF_TRACE(0, ti, 0);
                if (use > 0) mbar_wait(smem_u32(bar_empty + si), (use - 1) & 1);         // MMAs that read this stage retired
                CF_TRACE(0, ti, 1);
                const int b = tile / g.tiles_per_img, t = tile - b * g.tiles_per_img;
                const int r_lo = cf_row_lo(t * g.MT, g.halo, g.Wp);
                const uint32_t full = smem_u32(bar_landed + si);
                mbar_expect_tx(full, tx_bytes);
#pragma unroll
                for (int pl = 0; pl < KC; ++pl)
                    tma_load_4d(a_s + si * stage_bytes + (uint32_t)(pl * g.P) * 128, &tmap_x, full, pl * 64, -g.halo, r_lo, b);
                if (++si == S) { si = 0; ++use; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: the whole warp runs the (warp-uniform) loop, one elected lane issues =================
        constexpr uint32_t idesc = cf_idesc(COUT);
        uint64_t* bar_in = xform ? bar_ready : bar_landed;
        const uint64_t desc_hi = make_sw128_desc_bo(0, 0);              // SWIZZLE_128B K-major, SBO 1024, start address 0
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t leader = elect_one();
        const uint32_t w_lo = w_s >> 4;
        const int plane_a8 = g.P * 8;
        int toff[NTAPS];
#pragma unroll
        for (int t = 0; t < NTAPS; ++t) toff[t] = g.tap_off[t] * 8;     // a row of 128 B = 8 descriptor units of 16 B
        int si = 0, acc = 0, ti = 0;
        uint32_t in_phase = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++ti) {
            if (lane == 0) CF_TRACE(1, ti, 0);
            mbar_wait(smem_u32(bar_in + si), in_phase);                // tile staged (and transformed)
            if (lane == 0) CF_TRACE(1, ti, 1);
            fence_proxy_async_smem();
            tc_fence_after();
            const int t = tile % g.tiles_per_img;
            const int q0 = t * g.MT;
            const int pbase = q0 - cf_row_lo(q0, g.halo, g.Wp) * g.Wp;  // staged index of output position q0
            // descriptors differ only in their 14-bit start-address field (address >> 4): a K=16 step of 32 B = 2 units
            const int a_lo0 = (int)((a_s + si * stage_bytes) >> 4) + pbase * 8;
            for (int mm = 0; mm < MM; ++mm) {
                mbar_wait(smem_u32(bar_tempty + acc), acc_phase ^ 1);  // epilogue drained this accumulator
                if (lane == 0) CF_TRACE(1, ti, 2 + 2 * mm);
                tc_fence_after();
                const uint32_t d_tmem = tmem_u + acc * COUT;
                const int a_lo1 = a_lo0 + mm * 128 * 8;
#pragma unroll
                for (int tp = 0; tp < NTAPS; ++tp) {
#pragma unroll
                    for (int pl = 0; pl < KC; ++pl) {
#pragma unroll
                        for (int kk = 0; kk < KPP; ++kk) {
                            const uint32_t a = (uint32_t)(a_lo1 + toff[tp] + pl * plane_a8 + kk * 2);
                            const uint32_t b = w_lo + (uint32_t)((tp * KC + pl) * COUT * 8 + kk * 2);
                            umma_bf16_elect(leader, d_tmem, desc_hi | (uint64_t)a, desc_hi | (uint64_t)b, idesc, (tp | pl | kk) ? 1u : 0u);
                        }
                    }
                }
                umma_commit_elect(leader, smem_u32(bar_tfull + acc));  // accumulator complete -> epilogue
                if (lane == 0) CF_TRACE(1, ti, 3 + 2 * mm);
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
            umma_commit_elect(leader, smem_u32(bar_empty + si));        // stage free once these MMAs retire
            if (++si == S) { si = 0; in_phase ^= 1; }
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
                uint8_t* base = smem + W_BYTES + (size_t)si * stage_bytes + (size_t)((ch >> 3) * g.P) * 128;
                for (int rr = 0; rr < g.NR; ++rr) {
                    const int r = r_lo + rr;
                    if (r < 0 || r >= g.H) continue;                    // padding rows stay exactly zero
                    const int p0 = rr * g.Wp + g.halo;                  // interior columns only: padding columns stay zero
                    for (int c = c_first; c < g.W; c += PSTEP) {
                        const int p = p0 + c;
                        uint4* ptr = reinterpret_cast<uint4*>(base + (size_t)p * 128 + (((ch & 7) ^ (p & 7)) << 4));
                        float v[8];
                        cf_unpack(*ptr, v);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            v[k] = fmaf(v[k], sc[k], sh[k]);
                            if (relu) v[k] = fmaxf(v[k], 0.f);
                        }
                        store8(reinterpret_cast<__nv_bfloat16*>(ptr), v);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(bar_ready + si));
                if (ht == 0) CF_TRACE(2, ti, 2);
                if (++si == S) { si = 0; phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue: TMEM -> (+add, mask, reductions) -> bf16 -> global =================
        const int ew = warp - 6, grp = ew >> 2;                // group 0: channels [0, NC), group 1: [NC, COUT)
        const int q4 = warp & 3;                               // TMEM lane quarter this warp may access
        const int m = q4 * 32 + lane;
        const int ch0 = grp * NC;
        constexpr int NS = MODE == kCfPlain ? 1 : NC;
        float s1[NS], s2[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
        const bool has_add = ep.add != nullptr;
        const bool has_out = MODE == kCfBnRed && ep.bn_out != nullptr;
        const bool bn_relu = MODE == kCfBnRed && ep.bn_relu != 0;
        int acc = 0, ti = 0;
        uint32_t acc_phase = 0;
        const bool tracer = threadIdx.x == 192;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++ti) {
            const int b = tile / g.tiles_per_img, t = tile - b * g.tiles_per_img;
            for (int mm = 0; mm < MM; ++mm) {
                if (tracer) CF_TRACE(3, ti, 4 * mm);
                const int q = t * g.MT + mm * 128 + m;
                const int r = q / g.Wp, c = q - r * g.Wp - g.halo;
                const bool live = q < g.Q && c >= 0 && c < g.W;
                const size_t off = live ? (((size_t)b * g.H + r) * g.W + c) * COUT + ch0 : 0;
                uint4 av[NC / 8], zv[NC / 8], ov[NC / 8];
                if (live) {                                    // operands of the fused epilogue: in flight while the MMAs finish
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
                if (tracer) CF_TRACE(3, ti, 4 * mm + 1);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * COUT + ch0;
                uint32_t rr[NC];
#pragma unroll
                for (int c0 = 0; c0 < NC; c0 += 16) tmem_ld16(t_row + c0, rr + c0);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(bar_tempty + acc));      // the accumulator is in registers: hand TMEM back
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
                if (tracer) CF_TRACE(3, ti, 4 * mm + 2);
                if (!live) continue;
                __nv_bfloat16* dst = y + off;
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
                            const float d = v[k] - cst[ch0 + i * 8 + k];
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
                            bool on = true;
                            if (has_out) on = o8[k] > 0.f;
                            else if (bn_relu) on = fmaf(z8[k], cst[2 * COUT + cc], cst[3 * COUT + cc]) > 0.f;
                            const float gk = on ? v[k] : 0.f;
                            const float xh = (z8[k] - cst[cc]) * cst[COUT + cc];
                            s1[i * 8 + k] += gk;
                            s2[i * 8 + k] = fmaf(gk, xh, s2[i * 8 + k]);
                            v[k] = gk;
                        }
                    }
                    store8(dst + i * 8, v);
                }
                if (tracer) CF_TRACE(3, ti, 4 * mm + 3);
            }
        }
        if (MODE != kCfPlain) {
            // per-channel totals: warp shuffle tree -> 8 warp partials in smem -> one atomicAdd per channel per CTA
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const float a = warp_sum(s1[i]), b2 = warp_sum(s2[i]);
                if (lane == 0) { red[ew * 2 * NC + i] = a; red[ew * 2 * NC + NC + i] = b2; }
            }
            cf_epi_barrier();
            const int et = threadIdx.x - 192;                  // 0..255 within the epilogue group
            for (int i = et; i < 2 * COUT; i += kCfEpiThreads) {
                const int which = i / COUT, chn = i - which * COUT, gg = chn / NC, cl = chn - gg * NC;
                const float* rp = red + (gg * 4) * 2 * NC + which * NC + cl;
                atomicAdd(ep.accum + i, rp[0] + rp[2 * NC] + rp[4 * NC] + rp[6 * NC]);
            }
            __threadfence();
            cf_epi_barrier();
            if (et == 0) is_last = (atomicAdd(ep.ticket, 1u) == gridDim.x - 1);
            cf_epi_barrier();
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
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

struct CfPlan { CfGeom g; size_t smem; int grid; };

static bool cf_shape_ok(int Cin, int Cout, int n_taps, int mode);

static int cf_plan(int B, int H, int W, int Cin, int Cout, int n_taps, const int* dy, const int* dx, CfPlan* pl) {
    if (B <= 0 || H <= 0 || W <= 0 || n_taps < 1 || n_taps > kCfMaxTaps) return RSS_ERR_SHAPE;
    CfGeom& g = pl->g;
    g.B = B; g.H = H; g.W = W; g.in_relu = 0; g.trace = nullptr;
    int halo = 0;
    for (int t = 0; t < n_taps; ++t) {
        const int a = dy[t] < 0 ? -dy[t] : dy[t], b = dx[t] < 0 ? -dx[t] : dx[t];
        if (a > halo) halo = a;
        if (b > halo) halo = b;
    }
    if (halo > 1) return RSS_ERR_SHAPE;
    g.halo = halo; g.Wp = W + 2 * halo; g.Q = H * g.Wp;
    if (g.Wp > 256) return RSS_ERR_SHAPE;                                     // TMA box dimension limit
    for (int t = 0; t < kCfMaxTaps; ++t) g.tap_off[t] = t < n_taps ? dy[t] * g.Wp + dx[t] : 0;
    const int KC = (Cin + 63) / 64;
    const size_t w_bytes = (size_t)n_taps * KC * Cout * 128;
    const size_t budget = 227 * 1024 - 8 * 1024 - 1024;                       // static shared memory (<= 6.5 KB) + manual alignment
    // two 128-row blocks per tile halve the halo over-fetch; needs 4 accumulators in TMEM and a tile count that still fills the GPU
    for (int mm = (4 * Cout <= 512) ? 2 : 1; mm >= 1; --mm) {
        g.MM = mm; g.MT = 128 * mm;
        g.tiles_per_img = (g.Q + g.MT - 1) / g.MT;
        if (mm == 2 && (int64_t)B * g.tiles_per_img < 2 * num_sms()) continue;
        const int L = g.MT + 2 * halo;                                        // padded-linear span a tile reads within its own rows
        g.NR = (L + g.Wp - 2) / g.Wp + 1 + 2 * halo;                          // rows that span can touch, + halo rows above and below
        if (g.NR > 256) continue;
        g.P = (g.NR * g.Wp + 7) & ~7;                                         // multiple of 8 rows: planes/stages stay 1024-aligned
        const size_t stage = (size_t)KC * g.P * 128;
        for (int s = kCfMaxStages; s >= 2; --s) {
            const size_t need = w_bytes + (size_t)s * stage;
            if (need <= budget) {
                g.S = s; pl->smem = need + 1024;
                g.n_tiles = B * g.tiles_per_img;
                pl->grid = g.n_tiles < num_sms() ? g.n_tiles : num_sms();
                return RSS_OK;
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
    conv_cf_kernel<CIN, COUT, NTAPS, MODE><<<pl.grid, kCfThreads, pl.smem, stream>>>(
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
    return cf_plan(B, H, W, Cin, Cout, n, dy, dx, &pl) == RSS_OK;
}

// y = E(conv(T(x)) + add) with w_packed = bf16 [tap][Cout][Cin] from rss_conv_pack_weights (forward or transposed pack), taps
// (dy,dx) within [-1,1].  in_scale/in_shift (fp32 [Cin], NULL = identity) and in_relu describe T; `e` (may be NULL) the epilogue.
extern "C" int rss_conv_cf(const void* x, const void* w_packed, void* y, int B, int H, int W, int Cin, int Cout,
                           int n_taps, const int* taps_dy, const int* taps_dx,
                           const float* in_scale, const float* in_shift, int in_relu,
                           const RssConvCfEpilogue* e, cudaStream_t stream) {
    const int mode = e ? e->mode : RSS_CF_PLAIN;
    if (mode != RSS_CF_PLAIN && mode != RSS_CF_STATS && mode != RSS_CF_BNRED) return RSS_ERR_SHAPE;
    if (!cf_shape_ok(Cin, Cout, n_taps, mode)) return RSS_ERR_SHAPE;
    CfPlan pl;
    int rc = cf_plan(B, H, W, Cin, Cout, n_taps, taps_dy, taps_dx, &pl);
    if (rc != RSS_OK) return rc;
    if ((in_scale == nullptr) != (in_shift == nullptr)) return RSS_ERR_SHAPE;
    if (((uintptr_t)x & 15) || ((uintptr_t)y & 15) || ((uintptr_t)w_packed & 15)) return RSS_ERR_SHAPE;
    pl.g.in_relu = in_relu;
    {
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
        cuuint32_t box[4] = {64, (cuuint32_t)g.Wp, (cuuint32_t)g.NR, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
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
