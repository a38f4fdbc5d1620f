// Fused tcgen05 convolution for the HRNet family (BasicBlock/Bottleneck 3x3 and 1x1 stride-1 convs, their data
// gradients): _hrnet_rssformer.py:209-287.  NHWC bf16 in/out, fp32 TMEM accumulation.
//
//     y = conv( T(x) ),  T = identity | relu(x*scale + shift)      (the BatchNorm+ReLU of the PREVIOUS layer, applied on load)
//     + per-channel batch statistics of y (the BatchNorm that FOLLOWS), finalised by the last CTA
//
// Why not the tap-shifted-TMA kernel of conv_igemm.cu: that one re-fetches the A tile from L2 once per tap (9x for a 3x3), and
// ncu shows it pinned at the L2->SM fabric limit (profiles/ncu_igemm_probe_r1.csv).  Here every input pixel is staged in
// shared memory ONCE per tile and all taps read it at shifted addresses:
//   * positions are linearised over the zero-padded image, q = row*(W+2) + col+1, so that tap (dy,dx) of output q is input
//     q + dy*(W+2) + dx: a pure address offset (outputs that fall on a padding column are computed and dropped, 2/(W+2) waste);
//   * the staged tile is one 128-byte row (64 channels) per position in the canonical 128B-swizzled K-major UMMA layout
//     (16-byte chunk c of position p lives at p*128 + ((c ^ (p & 7)) << 4)), so a tap shift is "start address += offset*128 B"
//     in the shared-memory descriptor -- no re-load, no im2col.  Measured on B200: the hardware applies the swizzle XOR to the
//     ABSOLUTE shared-memory address bits [7,10), so a start address on any 128-byte row works with base_offset = 0 (setting
//     base_offset = (start >> 7) & 7 produces garbage).  (A first version used the no-swizzle "interleave" layout, whose shifts need no
//     phase at all; it was bit-correct but its operand fetch ran at ~16 B/cycle: ~450 cycles per M128xN32xK16 MMA.)
//   * the tile is staged by ONE TMA box load per 64-channel plane: whole padded rows {64 ch, W+2 pixels from x=-1, NR rows from
//     y=r_lo} -- TMA's out-of-bounds zero fill IS the padding (columns -1 and W, rows -1 and H), and the dense box order
//     [row][x][64 ch] IS the padded-linear position order.  (A first version filled the tile with per-thread cp.async: its
//     address arithmetic alone cost ~11 k cycles per tile, profiles/ncu_cf_probe_r1.csv.)  Optionally 3 helper warps apply the
//     previous layer's BN+ReLU in place (padding kept zero), fence to the async proxy and hand the stage to the MMA warp;
//   * one thread issues tcgen05.mma (M=128 rows = 128 consecutive positions, N = Cout, K = 16 channels) for every tap; 4
//     epilogue warps drain TMEM (tcgen05.ld), round to bf16, store, and keep per-thread running sums of (y-K), (y-K)^2 per
//     channel for the whole persistent loop.
// HBM traffic: x read once (+halo rows through L2), y written once; the BN statistics pass and (optionally) the BN-apply
// pass of the previous layer disappear.
#include <stdlib.h>
#include "tc05.cuh"

namespace rss {

constexpr int kCfThreads = 288;          // warp 0 TMA producer, warps 1-3 input-transform helpers, warp 4 MMA issuer, warps 5-8 epilogue
constexpr int kCfHelpers = 96;
constexpr int kCfMaxTaps = 9;
constexpr int kCfMaxStages = 3;

struct CfGeom {
    int B, H, W, Cin, Cout;
    int halo, Wp, Q;                     // padded pitch W + 2*halo, positions per image H*Wp
    int MM, MT;                          // 128-row MMA blocks per tile, MT = 128*MM
    int tiles_per_img, n_tiles;
    int NR, P;                           // staged padded rows per tile; plane pitch in positions (>= NR*Wp, multiple of 8)
    int KC;                              // 64-channel planes per position: ceil(Cin/64)
    int S;                               // ring stages (2 or 3)
    int n_taps;
    int tap_off[kCfMaxTaps];             // dy*Wp + dx (signed)
    int in_relu;
    int desc_swap;                       // debugging aid (RSS_CF_DESC_SWAP=1): base-offset field = (start >> 7) & 7 (measured WRONG)
    int dbg;                             // profiling aid (RSS_CF_DBG bits): 1 no epilogue stores/stats, 2 no TMA after the first S tiles, 4 no MMAs
    long long* trace;                    // profiling aid (RSS_CF_TRACE_PTR): CTA 0 records clock64() per role/tile/event, [4 roles][16 tiles][8]
};
#define CF_TRACE(role, i, k) do { if (g.trace && blockIdx.x == 0 && (i) < 16) g.trace[((role) * 16 + (i)) * 8 + (k)] = clock64(); } while (0)

struct CfStats {                         // all NULL when no statistics are wanted (data gradients)
    float* accum;                        // [2*Cout] persistent, zero between launches
    unsigned int* ticket;                // persistent, zero between launches
    const float* gamma; const float* beta;
    float* running_mean; float* running_var;     // may be NULL
    float momentum, eps;
    float* mean_out; float* invstd_out; float* scale_out; float* shift_out;
    float count;                         // B*H*W
};

__device__ __forceinline__ uint32_t cf_idesc(int n) {       // kind::f16, D=f32, A=B=bf16, K-major both, M=128
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void cf_epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// first staged padded row of the tile starting at padded-linear position q0 (floor division, q0 - halo may be negative)
__device__ __forceinline__ int cf_row_lo(int q0, int halo, int Wp) { return (q0 - halo + Wp) / Wp - 1 - halo; }

// COUT_S: compile-time Cout when statistics are produced (32 or 64), 0 = no statistics (Cout from the geometry).
template <int COUT_S>
__global__ void __launch_bounds__(kCfThreads, 1)
conv_cf_kernel(const __grid_constant__ CUtensorMap tmap_x, const __nv_bfloat16* __restrict__ wp, __nv_bfloat16* __restrict__ y,
               const float* __restrict__ in_scale, const float* __restrict__ in_shift, const __grid_constant__ CfGeom g,
               const __grid_constant__ CfStats st) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int S = g.S;
    const int CH = g.Cin >> 3;                                        // 16-byte channel chunks per position
    // every operand tile starts 1024-byte aligned (one swizzle period); P % 8 == 0 and Cout % 8 == 0 keep it so
    const uint32_t w_bytes = (uint32_t)g.n_taps * g.KC * g.Cout * 128;         // [tap][plane][cout row of 128 B]
    const uint32_t stage_bytes = (uint32_t)g.KC * g.P * 128;                   // [plane][position row of 128 B]
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t w_s = smem_u32(smem);
    const uint32_t a_s = w_s + w_bytes;
    uint8_t* tail = smem + w_bytes + (size_t)S * stage_bytes;
    // barriers: [0,3) landed (TMA bytes), [3,6) ready (transformed), [6,9) empty, [9,13) tmem_full, [13,17) tmem_empty
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    uint64_t* bar_landed = bars, *bar_ready = bars + 3, *bar_empty = bars + 6, *bar_tfull = bars + 9, *bar_tempty = bars + 13;
    const int NACC = 2 * g.MM;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);                 // +144 B; `red` below lands 16-byte aligned
    float* red = reinterpret_cast<float*>(tmem_slot + 4);            // [4 warps][2*Cout] statistics staging, then K[Cout]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool xform = in_scale != nullptr;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < NACC * g.Cout) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kCfMaxStages; ++s) {
            mbar_init(smem_u32(bar_landed + s), 1); mbar_init(smem_u32(bar_ready + s), kCfHelpers / 32); mbar_init(smem_u32(bar_empty + s), 1);
        }
        for (int a = 0; a < 4; ++a) { mbar_init(smem_u32(bar_tfull + a), 1); mbar_init(smem_u32(bar_tempty + a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    }
    if (warp == 4) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    // weights: packed global [tap][co][ci] -> smem [tap][plane][co][64 ci] rows of 128 B, 128B-swizzled (K-major B operand)
    {
        const int total = g.n_taps * g.Cout * CH;
        const int chs = CH == 4 ? 2 : (CH == 8 ? 3 : 4);                    // CH in {4, 8, 16}
        for (int i = threadIdx.x; i < total; i += kCfThreads) {
            const int kc = i & (CH - 1), row = i >> chs, co = row % g.Cout, tap = row / g.Cout;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(wp + ((size_t)(tap * g.Cout + co) * g.Cin + kc * 8)));
            const int plane = kc >> 3, c = kc & 7;
            *reinterpret_cast<uint4*>(smem + ((size_t)(tap * g.KC + plane) * g.Cout + co) * 128 + ((c ^ (co & 7)) << 4)) = v;
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: NR padded rows x Wp pixels x 64 channels per plane, OOB zero fill = padding =================
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)g.KC * g.NR * g.Wp * 128;
            int i = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++i) {
                const int si = i % S, use = i / S;
                CF_TRACE(0, i, 0);
                if (use > 0) mbar_wait(smem_u32(bar_empty + si), (use - 1) & 1);         // MMAs that read this stage retired
                CF_TRACE(0, i, 1);
                const int b = tile / g.tiles_per_img, t = tile % g.tiles_per_img;
                const int r_lo = cf_row_lo(t * g.MT, g.halo, g.Wp);
                const uint32_t full = smem_u32(bar_landed + si);
                if ((g.dbg & 2) && use > 0) { mbar_arrive(full); continue; }
                mbar_expect_tx(full, tx_bytes);
                for (int pl = 0; pl < g.KC; ++pl)
                    tma_load_4d(a_s + si * stage_bytes + (uint32_t)(pl * g.P) * 128, &tmap_x, full, pl * 64, -g.halo, r_lo, b);
                CF_TRACE(0, i, 2);
            }
        }
    } else if (warp < 4) {
        // ================= helpers: previous layer's BN(+ReLU) applied in place on the landed tile =================
        if (xform) {
            const int ht = threadIdx.x - 32;                      // 0..95
            const int ch = ht % CH, pslot = ht / CH, step = kCfHelpers / CH;      // CH in {4, 8, 16} divides 96
            float sc[8], sh[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { sc[k] = in_scale[ch * 8 + k]; sh[k] = in_shift[ch * 8 + k]; }
            const int npos = g.NR * g.Wp;
            int i = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++i) {
                const int si = i % S;
                if (ht == 0) CF_TRACE(1, i, 0);
                mbar_wait(smem_u32(bar_landed + si), (i / S) & 1);
                if (ht == 0) CF_TRACE(1, i, 1);
                const int t = tile % g.tiles_per_img;
                const int r_lo = cf_row_lo(t * g.MT, g.halo, g.Wp);
                uint8_t* base = smem + w_bytes + (size_t)si * stage_bytes + (size_t)((ch >> 3) * g.P) * 128;
                int r = r_lo + pslot / g.Wp, c = pslot % g.Wp - g.halo;
                for (int p = pslot; p < npos; p += step) {
                    if (r >= 0 && r < g.H && c >= 0 && c < g.W) {               // padding stays exactly zero
                        uint4* ptr = reinterpret_cast<uint4*>(base + (size_t)p * 128 + (((ch & 7) ^ (p & 7)) << 4));
                        Raw8<__nv_bfloat16> raw;
                        raw.r = *ptr;
                        float v[8];
                        unpack8(raw, v);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            v[k] = fmaf(v[k], sc[k], sh[k]);
                            if (g.in_relu) v[k] = fmaxf(v[k], 0.f);
                        }
                        store8(reinterpret_cast<__nv_bfloat16*>(ptr), v);
                    }
                    c += step;
                    while (c >= g.Wp - g.halo) { c -= g.Wp; ++r; }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(bar_ready + si));
                if (ht == 0) CF_TRACE(1, i, 2);
            }
        }
    } else if (warp == 4) {
        // ================= MMA issuer: the whole warp runs the (warp-uniform) loop, one elected lane issues =================
        {
            const uint32_t idesc = cf_idesc(g.Cout);
            const int kpp = g.Cin >= 64 ? 4 : (g.Cin >> 4);                 // K=16 steps per 64-channel plane
            uint64_t* bar_in = xform ? bar_ready : bar_landed;
            const uint64_t desc_hi = make_sw128_desc_bo(0, 0);              // SWIZZLE_128B K-major, SBO 1024, start address 0
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t leader = elect_one();
            const uint32_t w_lo = w_s >> 4;
            const uint32_t plane_a8 = (uint32_t)g.P * 8, plane_b8 = (uint32_t)g.Cout * 8;
            int i = 0, acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++i) {
                const int si = i % S;
                if (lane == 0) CF_TRACE(2, i, 0);
                mbar_wait(smem_u32(bar_in + si), (i / S) & 1);             // tile staged (and transformed)
                if (lane == 0) CF_TRACE(2, i, 1);
                fence_proxy_async_smem();
                tc_fence_after();
                const int t = tile % g.tiles_per_img;
                const int q0 = t * g.MT;
                const int pbase = q0 - cf_row_lo(q0, g.halo, g.Wp) * g.Wp;          // staged index of output position q0
                // descriptors differ only in their 14-bit start-address field (address >> 4): everything below is adds on that field
                // (a row of 128 B = 8 units, a K=16 step of 32 B = 2 units); the issue loop is the critical path of this kernel
                const uint32_t a_lo0 = ((a_s + si * stage_bytes) >> 4) + (uint32_t)pbase * 8;
                for (int mm = 0; mm < g.MM; ++mm) {
                    mbar_wait(smem_u32(bar_tempty + acc), acc_phase ^ 1);                // epilogue drained this accumulator
                    if (lane == 0) CF_TRACE(2, i, 2 + 2 * mm);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_u + acc * g.Cout;
                    const uint32_t a_lo1 = a_lo0 + (uint32_t)mm * 128 * 8;
                    uint32_t b_lo = w_lo;
                    uint32_t accum = 0;
                    for (int tp = 0; tp < g.n_taps; ++tp) {
                        uint32_t a_lo = a_lo1 + (uint32_t)(g.tap_off[tp] * 8);
                        for (int pl = 0; pl < g.KC; ++pl) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                if (kk < kpp && !(g.dbg & 4)) {
                                    umma_bf16_elect(leader, d_tmem, desc_hi | (uint64_t)(a_lo + kk * 2), desc_hi | (uint64_t)(b_lo + kk * 2), idesc, accum);
                                    accum = 1;
                                }
                            }
                            a_lo += plane_a8;
                            b_lo += plane_b8;
                        }
                    }
                    umma_commit_elect(leader, smem_u32(bar_tfull + acc));                        // accumulator complete -> epilogue
                    if (lane == 0) CF_TRACE(2, i, 3 + 2 * mm);
                    if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
                }
                umma_commit_elect(leader, smem_u32(bar_empty + si));                              // stage free once these MMAs retire
            }
        }
    } else {
        // ================= epilogue: TMEM -> bf16 -> global, running BN statistics =================
        const int q4 = warp & 3;                              // TMEM lane quarter this warp may access
        const int m = q4 * 32 + lane;
        constexpr int NS = COUT_S > 0 ? COUT_S : 1;
        float s1[NS], s2[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
        float* Ksm = red + 4 * 2 * g.Cout;                    // [Cout] shift of the one-pass variance (running mean)
        if (COUT_S > 0) {
            for (int c = threadIdx.x - 160; c < g.Cout; c += 128) Ksm[c] = st.running_mean ? st.running_mean[c] : 0.f;
            cf_epi_barrier();
        }
        int acc = 0;
        uint32_t acc_phase = 0;
        int ti = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++ti) {
            const int b = tile / g.tiles_per_img, t = tile % g.tiles_per_img;
            for (int mm = 0; mm < g.MM; ++mm) {
                const int q = t * g.MT + mm * 128 + m;
                const int r = q / g.Wp, c = q - r * g.Wp - g.halo;
                const bool live = q < g.Q && c >= 0 && c < g.W && !(g.dbg & 1);
                __nv_bfloat16* dst = y + (((size_t)b * g.H + r) * g.W + c) * g.Cout;
                if (m == 0) CF_TRACE(3, ti, 4 * mm);
                mbar_wait(smem_u32(bar_tfull + acc), acc_phase);
                if (m == 0) CF_TRACE(3, ti, 4 * mm + 1);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * g.Cout;
                if (COUT_S > 0) {
#pragma unroll
                    for (int c0 = 0; c0 < NS; c0 += 16) {
                        uint32_t rr[16];
                        tmem_ld16(t_row + c0, rr);
                        tmem_ld_wait();
                        if (live) {
                            float v[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[i]);
                            store8(dst + c0, v);
                            store8(dst + c0 + 8, v + 8);
#pragma unroll
                            for (int i4 = 0; i4 < 16; i4 += 4) {          // statistics of the fp32 accumulators (rounding is zero-mean)
                                const float4 k4 = *reinterpret_cast<const float4*>(Ksm + c0 + i4);
                                const float d0 = v[i4] - k4.x, d1 = v[i4 + 1] - k4.y, d2 = v[i4 + 2] - k4.z, d3 = v[i4 + 3] - k4.w;
                                s1[c0 + i4] += d0; s1[c0 + i4 + 1] += d1; s1[c0 + i4 + 2] += d2; s1[c0 + i4 + 3] += d3;
                                s2[c0 + i4] = fmaf(d0, d0, s2[c0 + i4]); s2[c0 + i4 + 1] = fmaf(d1, d1, s2[c0 + i4 + 1]);
                                s2[c0 + i4 + 2] = fmaf(d2, d2, s2[c0 + i4 + 2]); s2[c0 + i4 + 3] = fmaf(d3, d3, s2[c0 + i4 + 3]);
                            }
                        }
                    }
                } else {
                    for (int c0 = 0; c0 < g.Cout; c0 += 16) {
                        uint32_t rr[16];
                        tmem_ld16(t_row + c0, rr);
                        tmem_ld_wait();
                        if (live) {
                            float v[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[i]);
                            store8(dst + c0, v);
                            store8(dst + c0 + 8, v + 8);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(bar_tempty + acc));
                if (m == 0) CF_TRACE(3, ti, 4 * mm + 2);
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
        if (COUT_S > 0) {
            // per-channel totals: warp shuffle tree -> 4 warp partials in smem -> one atomicAdd per channel per CTA
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const float a = warp_sum(s1[i]), b2 = warp_sum(s2[i]);
                if (lane == 0) { red[q4 * 2 * NS + i] = a; red[q4 * 2 * NS + NS + i] = b2; }
            }
            cf_epi_barrier();
            const int et = threadIdx.x - 160;                  // 0..127 within the epilogue group
            for (int i = et; i < 2 * NS; i += 128)
                atomicAdd(st.accum + i, red[i] + red[2 * NS + i] + red[4 * NS + i] + red[6 * NS + i]);
            __threadfence();
            cf_epi_barrier();
            __shared__ bool is_last;
            if (et == 0) is_last = (atomicAdd(st.ticket, 1u) == gridDim.x - 1);
            cf_epi_barrier();
            if (is_last) {
                __threadfence();
                for (int c = et; c < NS; c += 128) {
                    const float Ssum = __ldcg(st.accum + c), Qsum = __ldcg(st.accum + NS + c);
                    const float md = Ssum / st.count;
                    const float m2 = fmaxf(Qsum - Ssum * md, 0.f);
                    const float mean = Ksm[c] + md;
                    const float invstd = rsqrtf(m2 / st.count + st.eps);
                    st.mean_out[c] = mean;
                    st.invstd_out[c] = invstd;
                    const float scl = st.gamma[c] * invstd;
                    st.scale_out[c] = scl;
                    st.shift_out[c] = st.beta[c] - mean * scl;
                    if (st.running_mean) {
                        st.running_mean[c] = (1.f - st.momentum) * st.running_mean[c] + st.momentum * mean;
                        st.running_var[c] = (1.f - st.momentum) * st.running_var[c] + st.momentum * (m2 / fmaxf(st.count - 1.f, 1.f));
                    }
                    st.accum[c] = 0.f;
                    st.accum[NS + c] = 0.f;
                }
                if (et == 0) *st.ticket = 0u;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, tmem_cols);
}

struct CfPlan { CfGeom g; size_t smem; int grid; };

static int cf_plan(int B, int H, int W, int Cin, int Cout, int n_taps, const int* dy, const int* dx, CfPlan* pl) {
    if (B <= 0 || H <= 0 || W <= 0 || n_taps < 1 || n_taps > kCfMaxTaps) return RSS_ERR_SHAPE;
    if (Cin != 32 && Cin != 64 && Cin != 128) return RSS_ERR_SHAPE;          // 96 helper threads / (Cin/8) chunks; 64-channel planes
    if (Cout < 16 || Cout % 16 || Cout > 128) return RSS_ERR_SHAPE;
    CfGeom& g = pl->g;
    g.B = B; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.n_taps = n_taps; g.in_relu = 0; g.desc_swap = 0; g.dbg = 0; g.trace = nullptr;
    int halo = 0;
    for (int t = 0; t < n_taps; ++t) {
        const int a = dy[t] < 0 ? -dy[t] : dy[t], b = dx[t] < 0 ? -dx[t] : dx[t];
        if (a > halo) halo = a;
        if (b > halo) halo = b;
    }
    if (halo > 1) return RSS_ERR_SHAPE;
    g.halo = halo; g.Wp = W + 2 * halo; g.Q = H * g.Wp;
    if (g.Wp > 256) return RSS_ERR_SHAPE;                                     // TMA box dimension limit
    for (int t = 0; t < n_taps; ++t) g.tap_off[t] = dy[t] * g.Wp + dx[t];
    g.KC = (Cin + 63) / 64;
    const size_t w_bytes = (size_t)n_taps * g.KC * Cout * 128;
    const size_t tail = 18 * 8 + 16 + (size_t)(4 * 2 + 1) * Cout * 4 + 64;
    const size_t budget = 225 * 1024 - 1024;                                  // 1024: manual alignment of the dynamic segment
    // two 128-row blocks per tile halve the halo over-fetch on wide images; needs 4 accumulators in TMEM
    for (int mm = (W >= 128 && 4 * Cout <= 512) ? 2 : 1; mm >= 1; --mm) {
        g.MM = mm; g.MT = 128 * mm;
        const int L = g.MT + 2 * halo;                                        // padded-linear span a tile reads within its own rows
        g.NR = (L + g.Wp - 2) / g.Wp + 1 + 2 * halo;                          // rows that span can touch, + halo rows above and below
        if (g.NR > 256) continue;
        g.P = (g.NR * g.Wp + 7) & ~7;                                         // multiple of 8 rows: planes/stages stay 1024-aligned
        const size_t stage = (size_t)g.KC * g.P * 128;
        for (int s = kCfMaxStages; s >= 2; --s) {
            const size_t need = w_bytes + (size_t)s * stage + tail;
            if (need <= budget) {
                g.S = s; pl->smem = need + 1024;
                g.tiles_per_img = (g.Q + g.MT - 1) / g.MT;
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

// conv_c32.cu
int conv_c32_launch(const void* x, const void* w_packed, void* y, int B, int H, int W, const int* taps_dy, const int* taps_dx,
                    const float* in_scale, const float* in_shift, int in_relu,
                    float* stat_accum, unsigned int* stat_ticket, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps,
                    float* mean_out, float* invstd_out, float* scale_out, float* shift_out, cudaStream_t stream);

}  // namespace rss

using namespace rss;

// 1 when rss_conv_cf accepts the geometry (stride 1, taps within a 3x3 neighbourhood, Cin in {32,64,128}, Cout % 16 == 0 <= 128,
// weights + 2 staged tiles fit in shared memory); with_stats additionally needs Cout in {32, 64}
extern "C" int rss_conv_cf_supported(int B, int H, int W, int Cin, int Cout, int ksize, int with_stats) {
    if (ksize != 1 && ksize != 3) return 0;
    if (with_stats && Cout != 32 && Cout != 64) return 0;
    int dy[9], dx[9], n = 0;
    for (int a = 0; a < ksize; ++a) for (int b = 0; b < ksize; ++b) { dy[n] = a - ksize / 2; dx[n] = b - ksize / 2; ++n; }
    CfPlan pl;
    return cf_plan(B, H, W, Cin, Cout, n, dy, dx, &pl) == RSS_OK;
}

// y = conv(T(x)) with w_packed = bf16 [tap][Cout][Cin] from rss_conv_pack_weights (forward or transposed pack), taps (dy,dx) within
// [-1,1].  in_scale/in_shift (fp32 [Cin], NULL = identity) and in_relu describe T.  stat_accum != NULL additionally produces the
// training-mode BatchNorm statistics of y exactly like rss_bn_stats_fused (same persistent scratch contract).
extern "C" int rss_conv_cf(const void* x, const void* w_packed, void* y, int B, int H, int W, int Cin, int Cout,
                           int n_taps, const int* taps_dy, const int* taps_dx,
                           const float* in_scale, const float* in_shift, int in_relu,
                           float* stat_accum, unsigned int* stat_ticket, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, float momentum, float eps,
                           float* mean_out, float* invstd_out, float* scale_out, float* shift_out, cudaStream_t stream) {
    CfPlan pl;
    int rc = cf_plan(B, H, W, Cin, Cout, n_taps, taps_dy, taps_dx, &pl);
    if (rc != RSS_OK) return rc;
    if ((in_scale == nullptr) != (in_shift == nullptr)) return RSS_ERR_SHAPE;
    if ((uintptr_t)x & 15) return RSS_ERR_SHAPE;
    pl.g.in_relu = in_relu;
    {
        const char* sw = getenv("RSS_CF_DESC_SWAP");
        pl.g.desc_swap = (sw && sw[0] == '1') ? 1 : 0;
        const char* dbg = getenv("RSS_CF_DBG");
        pl.g.dbg = dbg ? atoi(dbg) : 0;
        const char* tr = getenv("RSS_CF_TRACE_PTR");
        pl.g.trace = tr ? (long long*)strtoull(tr, nullptr, 16) : nullptr;
    }
    CfStats st{};
    const bool stats = stat_accum != nullptr;
    if (stats) {
        if ((Cout != 32 && Cout != 64) || !stat_ticket || !gamma || !beta || !mean_out || !invstd_out || !scale_out || !shift_out)
            return RSS_ERR_SHAPE;
        st.accum = stat_accum; st.ticket = stat_ticket; st.gamma = gamma; st.beta = beta; st.running_mean = running_mean;
        st.running_var = running_var; st.momentum = momentum; st.eps = eps; st.mean_out = mean_out; st.invstd_out = invstd_out;
        st.scale_out = scale_out; st.shift_out = shift_out; st.count = (float)((double)B * H * W);
    }
    {
        // EXPERIMENTAL (RSS_CF_MMA=1): the 32 -> 32 channel 3x3 layers go to the flat mma.sync kernel of conv_c32.cu
        const char* um = getenv("RSS_CF_MMA"); const int use_mma = um ? atoi(um) : 0;
        if (use_mma && Cin == 32 && Cout == 32 && n_taps == 9)
            return conv_c32_launch(x, w_packed, y, B, H, W, taps_dy, taps_dx, in_scale, in_shift, in_relu,
                                   stats ? stat_accum : nullptr, stat_ticket, gamma, beta, running_mean, running_var, momentum, eps,
                                   mean_out, invstd_out, scale_out, shift_out, stream);
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
    cudaError_t e = cudaSuccess;
#define CF_LAUNCH(CS)                                                                                                       \
    do {                                                                                                                     \
        static bool attr = false;                                                                                            \
        if (!attr) {                                                                                                         \
            e = cudaFuncSetAttribute(conv_cf_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(225 * 1024));   \
            attr = (e == cudaSuccess);                                                                                       \
        }                                                                                                                    \
        if (e == cudaSuccess)                                                                                                \
            conv_cf_kernel<CS><<<pl.grid, kCfThreads, pl.smem, stream>>>(tm, (const __nv_bfloat16*)w_packed, (__nv_bfloat16*)y, \
                                                                         in_scale, in_shift, pl.g, st);                     \
    } while (0)
    const int cs = stats ? Cout : 0;
    if (cs == 0) CF_LAUNCH(0);
    else if (cs == 32) CF_LAUNCH(32);
    else CF_LAUNCH(64);
#undef CF_LAUNCH
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; (void)cudaGetLastError(); return RSS_ERR_CUDA; }
    return check_launch();
}
