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
//   * 4 producer warps fill the tile with cp.async (zero fill = padding), optionally apply the previous layer's BN+ReLU in
//     place, fence to the async proxy and arrive on the stage's mbarrier; one thread issues tcgen05.mma (M=128 rows = 128
//     consecutive positions, N = Cout, K = 16 channels) for every tap; 4 epilogue warps drain TMEM (tcgen05.ld), round to
//     bf16, store, and keep per-thread running sums of (y-K), (y-K)^2 per channel for the whole persistent loop.
// HBM traffic: x read once (+halo rows through L2), y written once; the BN statistics pass and (optionally) the BN-apply
// pass of the previous layer disappear.
#include <stdlib.h>
#include "tc05.cuh"

namespace rss {

constexpr int kCfThreads = 288;          // warps 0-3 producers, warp 4 MMA issuer, warps 5-8 epilogue
constexpr int kCfProducers = 128;
constexpr int kCfMaxTaps = 9;

struct CfGeom {
    int B, H, W, Cin, Cout;
    int halo, Wp, Q;                     // padded pitch W + 2*halo, positions per image H*Wp
    int MM, MT;                          // 128-row MMA blocks per tile, MT = 128*MM
    int tiles_per_img, n_tiles;
    int P;                               // staged positions per tile: MT + 2*halo*(Wp+1)
    int n_taps;
    int tap_off[kCfMaxTaps];             // (dy+halo)*Wp + dx + halo
    int in_relu;
    int KC;                              // 64-channel planes per position: ceil(Cin/64)
    int desc_swap;                       // debugging aid (RSS_CF_DESC_SWAP=1): base-offset field = (start >> 7) & 7 (measured WRONG)
};

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
__device__ __forceinline__ void cf_cp16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cf_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cf_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cf_epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// COUT_S: compile-time Cout when statistics are produced (32 or 64), 0 = no statistics (Cout from the geometry).
// LA: producer look-ahead in tiles; the ring has LA+1 stages.
template <int COUT_S, int LA>
__global__ void __launch_bounds__(kCfThreads, 1)
conv_cf_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ wp, __nv_bfloat16* __restrict__ y,
               const float* __restrict__ in_scale, const float* __restrict__ in_shift, const __grid_constant__ CfGeom g,
               const __grid_constant__ CfStats st) {
    constexpr int S = LA + 1;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int CH = g.Cin >> 3;                                        // 16-byte channel chunks per position
    // every operand tile starts 1024-byte aligned (one swizzle period); P % 8 == 0 and Cout % 8 == 0 keep it so
    const uint32_t w_bytes = (uint32_t)g.n_taps * g.KC * g.Cout * 128;         // [tap][plane][cout row of 128 B]
    const uint32_t stage_bytes = (uint32_t)g.KC * g.P * 128;                   // [plane][position row of 128 B]
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t w_s = smem_u32(smem);
    const uint32_t a_s = w_s + w_bytes;
    uint8_t* tail = smem + w_bytes + (size_t)S * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);              // [0,S) full, [S,2S) empty, [2S,2S+NACC) tmem_full, then tmem_empty
    const int NACC = 2 * g.MM;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 2 * 4);
    float* red = reinterpret_cast<float*>(tmem_slot + 4);            // [4 warps][2*Cout] statistics staging
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < NACC * g.Cout) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(smem_u32(bars + s), kCfProducers / 32); mbar_init(smem_u32(bars + S + s), 1); }
        for (int a = 0; a < NACC; ++a) { mbar_init(smem_u32(bars + 2 * S + a), 1); mbar_init(smem_u32(bars + 2 * S + 4 + a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    // weights: packed global [tap][co][ci] -> smem [tap][plane][co][64 ci] rows of 128 B, 128B-swizzled (K-major B operand)
    {
        const int total = g.n_taps * g.Cout * CH;
        for (int i = threadIdx.x; i < total; i += kCfThreads) {
            const int kc = i % CH, co = (i / CH) % g.Cout, tap = i / (CH * g.Cout);
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

    if (warp < 4) {
        // ================= producers: global -> [chunk][position][8ch] shared tile =================
        const int tid = threadIdx.x;
        const int ch = tid % CH, pslot = tid / CH, step = kCfProducers / CH;
        const bool xform = in_scale != nullptr;
        float sc[8], sh[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { sc[i] = xform ? in_scale[ch * 8 + i] : 1.f; sh[i] = xform ? in_shift[ch * 8 + i] : 0.f; }

        auto issue = [&](int tile, int stage) {
            const int b = tile / g.tiles_per_img, t = tile % g.tiles_per_img;
            const int qs = t * g.MT - g.halo * g.Wp - g.halo + pslot;        // padded linear index of this thread's first position
            int r = (qs + 2 * g.Wp) / g.Wp - 2;
            int c = qs - r * g.Wp - g.halo;
            const uint32_t dst0 = a_s + stage * stage_bytes + (uint32_t)((ch >> 3) * g.P) * 128;
            const __nv_bfloat16* img = x + (size_t)b * g.H * g.W * g.Cin + ch * 8;
            for (int p = pslot; p < g.P; p += step) {
                const bool ok = r >= 0 && r < g.H && c >= 0 && c < g.W;
                cf_cp16(dst0 + (uint32_t)p * 128 + (uint32_t)(((ch & 7) ^ (p & 7)) << 4), ok ? img + ((size_t)r * g.W + c) * g.Cin : x, ok);
                c += step;
                while (c >= g.Wp - g.halo) { c -= g.Wp; ++r; }
            }
        };
        auto transform = [&](int tile, int stage) {
            const int t = tile % g.tiles_per_img;
            const int qs = t * g.MT - g.halo * g.Wp - g.halo + pslot;
            int r = (qs + 2 * g.Wp) / g.Wp - 2;
            int c = qs - r * g.Wp - g.halo;
            uint8_t* dst0 = smem + (a_s - w_s) + (size_t)stage * stage_bytes + (size_t)((ch >> 3) * g.P) * 128;
            for (int p = pslot; p < g.P; p += step) {
                if (r >= 0 && r < g.H && c >= 0 && c < g.W) {               // padding stays exactly zero
                    uint4* ptr = reinterpret_cast<uint4*>(dst0 + (size_t)p * 128 + (((ch & 7) ^ (p & 7)) << 4));
                    Raw8<__nv_bfloat16> raw;
                    raw.r = *ptr;
                    float v[8];
                    unpack8(raw, v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v[i] = fmaf(v[i], sc[i], sh[i]);
                        if (g.in_relu) v[i] = fmaxf(v[i], 0.f);
                    }
                    store8(reinterpret_cast<__nv_bfloat16*>(ptr), v);
                }
                c += step;
                while (c >= g.Wp - g.halo) { c -= g.Wp; ++r; }
            }
        };

        // ring bookkeeping: tile number i of this CTA uses stage i % S; a stage is re-filled only after the MMAs that read it retired
        int n_my = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) ++n_my;
        for (int i = 0; i < LA; ++i) {                                     // prologue: the first LA tiles (stages are free)
            if (i < n_my) issue(blockIdx.x + i * gridDim.x, i % S);
            cf_commit();
        }
        for (int i = 0; i < n_my; ++i) {
            const int j = i + LA;                                          // tile to prefetch now
            if (j < n_my) {
                const int sj = j % S;
                if (j >= S) mbar_wait(smem_u32(bars + S + sj), ((j / S) - 1) & 1);   // MMAs of tile j-S done with this stage
                issue(blockIdx.x + j * gridDim.x, sj);
            }
            cf_commit();
            cf_wait<LA>();                                                 // this thread's copies of tile i have landed
            const int si = i % S;
            if (xform) transform(blockIdx.x + i * gridDim.x, si);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(bars + si));
        }
        cf_wait<0>();
    } else if (warp == 4) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = cf_idesc(g.Cout);
            const int ksteps = g.Cin >> 4;
            int i = 0, acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++i) {
                const int si = i % S;
                mbar_wait(smem_u32(bars + si), (i / S) & 1);               // tile staged
                fence_proxy_async_smem();
                tc_fence_after();
                const uint32_t a0 = a_s + si * stage_bytes;
                for (int mm = 0; mm < g.MM; ++mm) {
                    mbar_wait(smem_u32(bars + 2 * S + 4 + acc), acc_phase ^ 1);          // epilogue drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * g.Cout;
                    for (int t = 0; t < g.n_taps; ++t) {
                        for (int k = 0; k < ksteps; ++k) {
                            const int plane = k >> 2, kk = k & 3;                        // 4 K=16 steps (32 B each) per 128-byte row
                            const uint32_t a_addr = a0 + (uint32_t)(plane * g.P + mm * 128 + g.tap_off[t]) * 128 + kk * 32;
                            const uint32_t b_addr = w_s + (uint32_t)((t * g.KC + plane) * g.Cout) * 128 + kk * 32;
                            const uint32_t phase = g.desc_swap ? ((a_addr >> 7) & 7u) : 0u;    // measured: the XOR uses absolute address bits
                            umma_bf16(d_tmem, make_sw128_desc_bo(a_addr, phase), make_sw128_desc_bo(b_addr, 0), idesc, (t | k) != 0);
                        }
                    }
                    umma_commit(smem_u32(bars + 2 * S + acc));                           // accumulator complete -> epilogue
                    if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
                }
                umma_commit(smem_u32(bars + S + si));                                     // stage free once these MMAs retire
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
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
            const int b = tile / g.tiles_per_img, t = tile % g.tiles_per_img;
            for (int mm = 0; mm < g.MM; ++mm) {
                const int q = t * g.MT + mm * 128 + m;
                const int r = q / g.Wp, c = q - r * g.Wp - g.halo;
                const bool live = q < g.Q && c >= 0 && c < g.W;
                __nv_bfloat16* dst = y + (((size_t)b * g.H + r) * g.W + c) * g.Cout;
                mbar_wait(smem_u32(bars + 2 * S + acc), acc_phase);
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
                            for (int i = 0; i < 16; ++i) {                 // statistics of the ROUNDED values (what the apply pass reads)
                                const float d = __bfloat162float(__float2bfloat16_rn(v[i])) - Ksm[c0 + i];
                                s1[c0 + i] += d;
                                s2[c0 + i] = fmaf(d, d, s2[c0 + i]);
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
                if (lane == 0) mbar_arrive(smem_u32(bars + 2 * S + 4 + acc));
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

struct CfPlan { CfGeom g; size_t smem; int la; int grid; };

static int cf_plan(int B, int H, int W, int Cin, int Cout, int n_taps, const int* dy, const int* dx, CfPlan* pl) {
    if (B <= 0 || H <= 0 || W <= 0 || n_taps < 1 || n_taps > kCfMaxTaps) return RSS_ERR_SHAPE;
    if (Cin != 32 && Cin != 64 && Cin != 128) return RSS_ERR_SHAPE;          // 128 producer threads / (Cin/8) chunks
    if (Cout < 16 || Cout % 16 || Cout > 128) return RSS_ERR_SHAPE;
    CfGeom& g = pl->g;
    g.B = B; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.n_taps = n_taps; g.in_relu = 0;
    int halo = 0;
    for (int t = 0; t < n_taps; ++t) {
        const int a = dy[t] < 0 ? -dy[t] : dy[t], b = dx[t] < 0 ? -dx[t] : dx[t];
        if (a > halo) halo = a;
        if (b > halo) halo = b;
    }
    if (halo > 1) return RSS_ERR_SHAPE;
    g.halo = halo; g.Wp = W + 2 * halo; g.Q = H * g.Wp;
    for (int t = 0; t < n_taps; ++t) g.tap_off[t] = (dy[t] + halo) * g.Wp + dx[t] + halo;
    g.KC = (Cin + 63) / 64;
    const size_t w_bytes = (size_t)n_taps * g.KC * Cout * 128;
    const size_t tail = (2 * 3 + 8) * 8 + 16 + (size_t)(4 * 2 + 1) * Cout * 4 + 64;
    const size_t budget = 225 * 1024 - 1024;                                  // 1024: manual alignment of the dynamic segment
    // two 128-row blocks per tile halve the halo over-fetch on wide images; needs 4 accumulators in TMEM
    for (int mm = (W >= 128 && 4 * Cout <= 512) ? 2 : 1; mm >= 1; --mm) {
        g.MM = mm; g.MT = 128 * mm;
        g.P = (g.MT + 2 * halo * (g.Wp + 1) + 7) & ~7;                        // multiple of 8 rows: planes/stages stay 1024-aligned
        const size_t stage = (size_t)g.KC * g.P * 128;
        for (int la = 2; la >= 1; --la) {
            const size_t need = w_bytes + (size_t)(la + 1) * stage + tail;
            if (need <= budget) {
                pl->la = la; pl->smem = need + 1024;
                g.tiles_per_img = (g.Q + g.MT - 1) / g.MT;
                g.n_tiles = B * g.tiles_per_img;
                pl->grid = g.n_tiles < num_sms() ? g.n_tiles : num_sms();
                return RSS_OK;
            }
        }
    }
    return RSS_ERR_SHAPE;
}

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
    pl.g.in_relu = in_relu;
    {
        const char* sw = getenv("RSS_CF_DESC_SWAP");
        pl.g.desc_swap = (sw && sw[0] == '1') ? 1 : 0;
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
    cudaError_t e = cudaSuccess;
#define CF_LAUNCH(CS, LA_)                                                                                                  \
    do {                                                                                                                     \
        static bool attr = false;                                                                                            \
        if (!attr) {                                                                                                         \
            e = cudaFuncSetAttribute(conv_cf_kernel<CS, LA_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(225 * 1024)); \
            attr = (e == cudaSuccess);                                                                                       \
        }                                                                                                                    \
        if (e == cudaSuccess)                                                                                                \
            conv_cf_kernel<CS, LA_><<<pl.grid, kCfThreads, pl.smem, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w_packed, \
                                                                              (__nv_bfloat16*)y, in_scale, in_shift, pl.g, st); \
    } while (0)
    const int cs = stats ? Cout : 0;
    if (cs == 0 && pl.la == 2) CF_LAUNCH(0, 2);
    else if (cs == 0) CF_LAUNCH(0, 1);
    else if (cs == 32 && pl.la == 2) CF_LAUNCH(32, 2);
    else if (cs == 32) CF_LAUNCH(32, 1);
    else if (cs == 64 && pl.la == 2) CF_LAUNCH(64, 2);
    else CF_LAUNCH(64, 1);
#undef CF_LAUNCH
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; (void)cudaGetLastError(); return RSS_ERR_CUDA; }
    return check_launch();
}
