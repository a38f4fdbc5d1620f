// Direct 3x3 convolution for the 32-channel branch of HRNet (BasicBlock convs of branch 0 and their data gradients,
// _hrnet_rssformer.py:216-246): NHWC bf16 in/out, fp32 accumulation on mma.sync m16n8k16, same contract as rss_conv_cf
// (optional BN+ReLU of the previous layer applied to the input on load, optional training-mode BatchNorm statistics of the output
// finalised by the last CTA).
//
// STATUS: EXPERIMENTAL, written at the end of round 1 after the GPU budget was spent -- compiled, not yet run on a B200.  It is only
// reachable with RSS_CF_MMA=1 (rss_conv_cf then routes Cin = Cout = 32, 9-tap calls here); `RSS_CONV_CF=1 RSS_CF_MMA=1 pytest -k
// conv_cf` runs the existing parity cases (CF_CASES rows 1-3) through it.  Nothing in the default path uses it.
//
// Why: these layers are HBM/latency-bound (4.8 GFLOP, 33.5 MB per call) and sit on the critical stream 64 + 64 times per step.
// The library runs them with a 23-37 us `sm80_xmma` kernel plus a separate statistics pass (12 us) plus a weight re-layout; the
// tcgen05 kernel of conv_cf.cu needs ~30 us per launch for its barrier skeleton alone.  This kernel keeps
// the structure as flat as possible:
//   * one CTA (4 warps) computes a 4 x 32 pixel tile: the 6 x 34 x 32-channel halo tile (13 KB) is fetched ONCE with 16-byte
//     cp.async (zero fill outside the image = the padding), the 9 x 32 x 32 weights (18 KB) stay in shared memory for the CTA's
//     whole persistent loop, and every tap reads the same staged tile at a shifted pixel index -- no im2col, no re-fetch;
//   * both tiles use 64-byte rows (one pixel / one output channel) with the 16-byte chunk index XOR-ed with (row >> 1) & 3, so the
//     8 rows of every ldmatrix hit 8 different bank groups;
//   * warp w owns output row w: 2 (m16) x 4 (n8) accumulator tiles, 144 MMAs per tile;
//   * 5 CTAs per SM (32 KB shared memory, <= 102 registers) overlap one CTA's loads with the others' MMAs instead of an
//     in-CTA multi-stage pipeline;
//   * epilogue: per-channel sum / sum of squares of the fp32 accumulators in registers across the persistent loop (one shuffle tree
//     + 64 atomics per CTA at the very end), outputs staged through shared memory into 16-byte coalesced stores.
#include <stdlib.h>
#include "common.cuh"

namespace rss {

constexpr int kC32 = 32;
constexpr int kC32TH = 4, kC32TW = 32;                   // output tile
constexpr int kC32HH = kC32TH + 2, kC32HW = kC32TW + 2;  // halo tile
constexpr int kC32Threads = 32 * kC32TH;

struct C32Geom {
    int B, H, W;
    int tiles_x, tiles_y, n_tiles;
    int dy[9], dx[9];                                    // input offset of tap t (|.| <= 1)
    int in_relu;
};

struct C32Stats {                                        // same contract as CfStats (conv_cf.cu); accum == NULL: no statistics
    float* accum; unsigned int* ticket;
    const float* gamma; const float* beta;
    float* running_mean; float* running_var;
    float momentum, eps;
    float* mean_out; float* invstd_out; float* scale_out; float* shift_out;
    float count;
};

__device__ __forceinline__ uint32_t c32_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void c32_ldsm_x4(uint32_t r[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void c32_mma(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void c32_cp_async16(uint32_t dst, const void* src, int src_bytes) {       // src_bytes 0 -> 16 zero bytes
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void c32_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint32_t c32_pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of 16-byte chunk c of 64-byte row `row` (swizzled)
__device__ __forceinline__ uint32_t c32_off(int row, int c) { return (uint32_t)row * 64u + (uint32_t)((c ^ ((row >> 1) & 3)) << 4); }

template <bool STATS>
__global__ void __launch_bounds__(kC32Threads, 5)
conv_c32_mma_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ wp /*[9][32 n][32 k]*/,
                    __nv_bfloat16* __restrict__ y, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                    const C32Geom g, const C32Stats st) {
    __shared__ __align__(128) uint8_t in_sm[kC32HH * kC32HW * 64];          // halo tile, later the output staging of each warp
    __shared__ __align__(128) uint8_t w_sm[9 * kC32 * 64];
    __shared__ float sc_sm[kC32], sh_sm[kC32], k_sm[kC32];
    __shared__ float red[kC32TH][2 * kC32];
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
    const bool xform = in_scale != nullptr;

    // ---- once per CTA: weights (already bf16 [tap][n][k], k contiguous) -> swizzled rows; per-channel constants
    for (int i = tid; i < 9 * kC32 * 4; i += kC32Threads) {
        const int q = i >> 2, c = i & 3;                                    // q = tap*32 + n
        *reinterpret_cast<uint4*>(w_sm + c32_off(q, c)) = __ldg(reinterpret_cast<const uint4*>(wp + (size_t)q * kC32 + c * 8));
    }
    if (tid < kC32) {
        sc_sm[tid] = xform ? in_scale[tid] : 1.f;
        sh_sm[tid] = xform ? in_shift[tid] : 0.f;
        k_sm[tid] = (STATS && st.running_mean) ? st.running_mean[tid] : 0.f;
    }
    float s1[4][2], s2[4][2];                                               // running statistics: channels nt*8 + 2*tq + e
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { s1[nt][0] = s1[nt][1] = s2[nt][0] = s2[nt][1] = 0.f; }
    __syncthreads();
    const uint32_t in_u = c32_smem(in_sm), w_u = c32_smem(w_sm);

    for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const int tx = tile % g.tiles_x, ty = (tile / g.tiles_x) % g.tiles_y, b = tile / (g.tiles_x * g.tiles_y);
        const int x0 = tx * kC32TW, y0 = ty * kC32TH;
        // ---- halo tile: rows y0-1 .. y0+4, columns x0-1 .. x0+32, zero outside the image
        for (int i = tid; i < kC32HH * kC32HW * 4; i += kC32Threads) {
            const int p = i >> 2, c = i & 3, ry = p / kC32HW, rx = p - ry * kC32HW;
            const int gy = y0 - 1 + ry, gx = x0 - 1 + rx;
            const bool in = gy >= 0 && gy < g.H && gx >= 0 && gx < g.W;
            const __nv_bfloat16* src = in ? x + (((size_t)b * g.H + gy) * g.W + gx) * kC32 + c * 8 : x;
            c32_cp_async16(in_u + c32_off(p, c), src, in ? 16 : 0);
        }
        c32_cp_async_wait_all();
        if (xform) {                                                        // BN(+ReLU) of the previous layer, on this thread's own chunks
            for (int i = tid; i < kC32HH * kC32HW * 4; i += kC32Threads) {
                const int p = i >> 2, c = i & 3, ry = p / kC32HW, rx = p - ry * kC32HW;
                const int gy = y0 - 1 + ry, gx = x0 - 1 + rx;
                if (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) {           // padding stays exactly zero
                    uint4* ptr = reinterpret_cast<uint4*>(in_sm + c32_off(p, c));
                    Raw8<__nv_bfloat16> raw;
                    raw.r = *ptr;
                    float v[8];
                    unpack8(raw, v);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        v[k] = fmaf(v[k], sc_sm[c * 8 + k], sh_sm[c * 8 + k]);
                        if (g.in_relu) v[k] = fmaxf(v[k], 0.f);
                    }
                    *ptr = make_uint4(c32_pack(v[0], v[1]), c32_pack(v[2], v[3]), c32_pack(v[4], v[5]), c32_pack(v[6], v[7]));
                }
            }
        }
        __syncthreads();
        // ---- 9 taps x 2 k-steps: acc[mt][nt] += A(tap-shifted pixels of row `warp`) . W_tap
        float acc[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int prow = (warp + 1 + g.dy[t]) * kC32HW + 1 + g.dx[t];   // halo index of output pixel (row warp, column 0) for this tap
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t a[2][4], bw[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {                            // lanes 0-15: rows (pixels) 0-15, lanes 16-31: same rows, k + 8
                    const int p = prow + mt * 16 + (lane & 15);
                    c32_ldsm_x4(a[mt], in_u + c32_off(p, ks * 2 + (lane >> 4)));
                }
#pragma unroll
                for (int np = 0; np < 2; ++np) {                            // matrices: (n-tile 2np, k lo), (2np, k hi), (2np+1, k lo), (2np+1, k hi)
                    const int q = t * kC32 + (np * 2 + (lane >> 4)) * 8 + (lane & 7);
                    c32_ldsm_x4(bw[np], w_u + c32_off(q, ks * 2 + ((lane >> 3) & 1)));
                }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) c32_mma(acc[mt][nt], a[mt], bw[nt >> 1][(nt & 1) * 2], bw[nt >> 1][(nt & 1) * 2 + 1]);
            }
        }
        // ---- epilogue
        const int oy = y0 + warp;
        if (STATS && oy < g.H) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int ox = x0 + mt * 16 + gq + half * 8;
                    if (ox < g.W) {
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const float d = acc[mt][nt][half * 2 + e] - k_sm[nt * 8 + tq * 2 + e];
                                s1[nt][e] += d;
                                s2[nt][e] = fmaf(d, d, s2[nt][e]);
                            }
                    }
                }
        }
        __syncthreads();                                                    // every warp is done reading the halo tile
        {
            // stage this warp's 32 pixels x 32 channels (2 KB, warp-private region of the old halo tile) and store 16 bytes per lane
            const uint32_t wbase = (uint32_t)warp * (kC32TW * 64);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int px = mt * 16 + gq + half * 8;
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
                        *reinterpret_cast<uint32_t*>(in_sm + wbase + c32_off(px, nt) + tq * 4) =
                            c32_pack(acc[mt][nt][half * 2], acc[mt][nt][half * 2 + 1]);
                }
            __syncwarp();
            if (oy < g.H) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int px = j * 8 + (lane >> 2), c = lane & 3, ox = x0 + px;
                    if (ox < g.W)
                        *reinterpret_cast<uint4*>(y + (((size_t)b * g.H + oy) * g.W + ox) * kC32 + c * 8) =
                            *reinterpret_cast<const uint4*>(in_sm + wbase + c32_off(px, c));
                }
            }
        }
        __syncthreads();                                                    // staging consumed before the next tile's cp.async lands
    }
    if (!STATS) return;
    // ---- statistics: shuffle tree over the 8 row groups -> 4 warp partials -> one atomic per channel and CTA -> last CTA finalises
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float a = s1[nt][e], q = s2[nt][e];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
            if (gq == 0) { red[warp][nt * 8 + tq * 2 + e] = a; red[warp][kC32 + nt * 8 + tq * 2 + e] = q; }
        }
    __syncthreads();
    if (tid < 2 * kC32) atomicAdd(st.accum + tid, red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid]);
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(st.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tid < kC32) {
        const int c = tid;
        const float S = __ldcg(st.accum + c), Q = __ldcg(st.accum + kC32 + c);
        const float md = S / st.count;
        const float m2 = fmaxf(Q - S * md, 0.f);
        const float mean = k_sm[c] + md;
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
        st.accum[kC32 + c] = 0.f;
    }
    if (tid == 0) *st.ticket = 0u;
}

// host side, called from rss_conv_cf (conv_cf.cu) when RSS_CF_MMA=1 and the geometry is Cin = Cout = 32 with 9 taps
int conv_c32_launch(const void* x, const void* w_packed, void* y, int B, int H, int W, const int* taps_dy, const int* taps_dx,
                    const float* in_scale, const float* in_shift, int in_relu,
                    float* stat_accum, unsigned int* stat_ticket, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps,
                    float* mean_out, float* invstd_out, float* scale_out, float* shift_out, cudaStream_t stream) {
    C32Geom g;
    g.B = B; g.H = H; g.W = W; g.in_relu = in_relu;
    g.tiles_x = (W + kC32TW - 1) / kC32TW; g.tiles_y = (H + kC32TH - 1) / kC32TH;
    const int64_t nt = (int64_t)B * g.tiles_x * g.tiles_y;
    if (nt <= 0 || nt > 0x7fffffff) return RSS_ERR_SHAPE;
    g.n_tiles = (int)nt;
    for (int t = 0; t < 9; ++t) {
        if (taps_dy[t] < -1 || taps_dy[t] > 1 || taps_dx[t] < -1 || taps_dx[t] > 1) return RSS_ERR_SHAPE;
        g.dy[t] = taps_dy[t]; g.dx[t] = taps_dx[t];
    }
    if (((uintptr_t)x & 15) || ((uintptr_t)y & 15) || ((uintptr_t)w_packed & 15)) return RSS_ERR_SHAPE;
    C32Stats st{};
    st.accum = stat_accum; st.ticket = stat_ticket; st.gamma = gamma; st.beta = beta; st.running_mean = running_mean;
    st.running_var = running_var; st.momentum = momentum; st.eps = eps; st.mean_out = mean_out; st.invstd_out = invstd_out;
    st.scale_out = scale_out; st.shift_out = shift_out; st.count = (float)((int64_t)B * H * W);
    int grid = num_sms() * 5;
    if (grid > g.n_tiles) grid = g.n_tiles;
    if (stat_accum)
        conv_c32_mma_kernel<true><<<grid, kC32Threads, 0, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w_packed,
                                                                     (__nv_bfloat16*)y, in_scale, in_shift, g, st);
    else
        conv_c32_mma_kernel<false><<<grid, kC32Threads, 0, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w_packed,
                                                                      (__nv_bfloat16*)y, in_scale, in_shift, g, st);
    return check_launch();
}

}  // namespace rss
