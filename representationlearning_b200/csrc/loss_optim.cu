// Fused x4 bilinear up-sampling + foreground-aware focal cross-entropy, and the fused clip+SGD step.
//
// Reference (RSSFormer-TIP2023):
//   head upsample ........ module/baseline/hrnet_aux.py:80 (UpsamplingBilinear2d == align_corners=True)
//   SegmentationLossaux .. module/CGFL.py:201-227 ; softmax_focalloss CGFL.py:72-101
//   MCTransAuxLoss ....... losses/auxloss.py:257-305 (per-image unique() in a Python loop = B host syncs)
//   optimiser ............ configs/base/loveda.py:68-99 (SGD m=0.9 wd=1e-4, clip_grad_norm_ 35, poly LR)
//
// Closed form implemented here (pinned against the reference in oracle/gen_golden.py):
//   fg presence: onehot_b[0] = any(label<=0) ; onehot_b[1] = any(label>0)
//   l1_b  = sum_c 1/(1+exp|s_bc - onehot_bc|) / (2B)
//   CE    = mean over valid pixels of -log softmax(z)[label]
//   loss  = CE * sum_b (1-l1_b/7) * A_b / (n_valid + B),   A_b = sum over ALL pixels of (1 - p_t), t = label or 0 if ignored
//   d loss / d z = (softmax(z) - onehot(label)) * valid * [sum_b(...)/(n_valid+B)] / n_valid     (factor is no_grad)
// One pass over the full-resolution pixels computes everything, including the gradient w.r.t. the
// LOW-resolution logits (bilinear transpose accumulated in shared memory), so the (B,7,512,512)
// logits/probabilities are never written to HBM and no host synchronisation happens.
#include "common.cuh"

namespace rss {

constexpr int kNC = 7, kNCP = 8, kTile = 32, kFoot = 18;   // footprint of a 32-wide output tile in input pixels: ceil(31*(h-1)/(H-1))+2 <= 18 for scale>=2

__device__ __forceinline__ void bilinear_src_l(int o, float scale, int in, int& i0, int& i1, float& l1) {
    const float src = scale * (float)o;
    i0 = (int)src;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

// acc layout (floats): [0] ce_sum, [1] n_valid, [2..2+B) A_b, then ints: [2+B .. 2+2B) has_fg, [2+2B .. 2+3B) has_bg
//
// Gradient w.r.t. the low-resolution logits = transpose of the bilinear up-sampling applied to the per-pixel softmax gradient.
// The transpose is done as a GATHER in two separable passes over shared memory (x, then y): every low-resolution cell sums the
// <= 9 output columns (rows) that touch it.  (A first version scattered with shared-memory float atomics: sm_100 has no native
// shared fp32 add -- SASS showed ATOMS.CAST.SPIN loops, ~16-way contended -- and the kernel took 674 us for 37 MB of traffic.)
// 512 threads per 32 x 32 tile (two output rows per thread in the per-pixel phase): at 1024 threads x 45 registers only ONE tile was
// resident per SM and its three barriers + the thin gather passes ran unoverlapped (6 us per tile); two resident tiles overlap them.
constexpr int kLossThreads = kTile * kTile / 2;
__global__ void __launch_bounds__(kLossThreads, 2)
seg_loss_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ acc,
                    float* __restrict__ gdir /*[B][h][w][8], zeroed*/, int B, int h, int w, int H, int W,
                    float sy, float sx, int ignore_index) {
    __shared__ float G[kTile][kTile * kNC + 1];     // per-pixel d CE / d z (0 where ignored / outside the image); [row][col*7 + class], odd pitch
    __shared__ float T[kTile][kFoot][kNC];          // after the x pass: [output row][low-res column][class]
    __shared__ int nf[2];                           // low-resolution columns / rows this tile's footprint really covers (<= kFoot)
    __shared__ float red[3][kLossThreads / 32];
    __shared__ int flags[2];
    // per output column / row: footprint-relative low source index (kBig outside the image), weight of the low and of the high
    // source cell (the high weight is folded into the low one where align_corners clamps both onto the last cell)
    __shared__ int cx0[kTile], cy0[kTile];
    __shared__ float wx0[kTile], wx1[kTile], wy0[kTile], wy1[kTile];
    __shared__ int fx[kFoot + 2], fy[kFoot + 2];     // fx[r] = number of columns with cx0 < r: columns [fx[r], fx[r+1]) have cx0 == r
    constexpr int kBig = 1 << 20;
    const int tiles_x = (W + kTile - 1) / kTile;
    const int ty0 = (blockIdx.x / tiles_x) * kTile, tx0 = (blockIdx.x % tiles_x) * kTile;
    const int b = blockIdx.y;
    int fy0, fx0, tmp; float tl;
    bilinear_src_l(ty0, sy, h, fy0, tmp, tl);
    bilinear_src_l(tx0, sx, w, fx0, tmp, tl);
    if (threadIdx.x < 2) flags[threadIdx.x] = 0;
    __syncthreads();
    const int c = threadIdx.x % kTile, ox = tx0 + c;
    float ce = 0.f, nv = 0.f, A = 0.f;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        const int r = threadIdx.x / kTile + half * (kTile / 2), oy = ty0 + r;
        float gk[kNC];
#pragma unroll
        for (int k = 0; k < kNC; ++k) gk[k] = 0.f;
        int y0, y1, x0, x1; float ly, lx;
        bilinear_src_l(oy < H ? oy : H - 1, sy, h, y0, y1, ly);
        bilinear_src_l(ox < W ? ox : W - 1, sx, w, x0, x1, lx);
        if (r == 0) { cx0[c] = ox < W ? x0 - fx0 : kBig; wx0[c] = x1 == x0 ? 1.f : 1.f - lx; wx1[c] = x1 == x0 ? 0.f : lx; }
        if (c == 0) { cy0[r] = oy < H ? y0 - fy0 : kBig; wy0[r] = y1 == y0 ? 1.f : 1.f - ly; wy1[r] = y1 == y0 ? 0.f : ly; }
        if (oy < H && ox < W) {
            const float* base = logits + (int64_t)b * h * w * kNCP;
            const float4* p00 = reinterpret_cast<const float4*>(base + ((int64_t)y0 * w + x0) * kNCP);
            const float4* p01 = reinterpret_cast<const float4*>(base + ((int64_t)y0 * w + x1) * kNCP);
            const float4* p10 = reinterpret_cast<const float4*>(base + ((int64_t)y1 * w + x0) * kNCP);
            const float4* p11 = reinterpret_cast<const float4*>(base + ((int64_t)y1 * w + x1) * kNCP);
            float a[8], bq[8], cc[8], d[8];
            *reinterpret_cast<float4*>(a) = __ldg(p00); *reinterpret_cast<float4*>(a + 4) = __ldg(p00 + 1);
            *reinterpret_cast<float4*>(bq) = __ldg(p01); *reinterpret_cast<float4*>(bq + 4) = __ldg(p01 + 1);
            *reinterpret_cast<float4*>(cc) = __ldg(p10); *reinterpret_cast<float4*>(cc + 4) = __ldg(p10 + 1);
            *reinterpret_cast<float4*>(d) = __ldg(p11); *reinterpret_cast<float4*>(d + 4) = __ldg(p11 + 1);
            const float hy = 1.f - ly, hx = 1.f - lx;
            float z[kNC], mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < kNC; ++k) {
                z[k] = hy * (hx * a[k] + lx * bq[k]) + ly * (hx * cc[k] + lx * d[k]);
                mx = fmaxf(mx, z[k]);
            }
            float e[kNC], s = 0.f;
#pragma unroll
            for (int k = 0; k < kNC; ++k) { e[k] = expf(z[k] - mx); s += e[k]; }
            const float lse = mx + logf(s), inv_s = 1.f / s;
            const int64_t lbl = labels[((int64_t)b * H + oy) * W + ox];
            const bool valid = lbl != (int64_t)ignore_index;
            const int t = valid ? (int)lbl : 0;
            float zt = 0.f, et = 0.f;
#pragma unroll
            for (int k = 0; k < kNC; ++k) { zt = (k == t) ? z[k] : zt; et = (k == t) ? e[k] : et; }
            A += 1.f - et * inv_s;
            if (lbl > 0) flags[0] = 1; else flags[1] = 1;          // benign race: all writers store 1
            if (valid) {
                ce += lse - zt;
                nv += 1.f;
#pragma unroll
                for (int k = 0; k < kNC; ++k) gk[k] = e[k] * inv_s - (k == t ? 1.f : 0.f);
            }
        }
#pragma unroll
        for (int k = 0; k < kNC; ++k) G[r][c * kNC + k] = gk[k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ce = warp_sum(ce); nv = warp_sum(nv); A = warp_sum(A);
    if (lane == 0) { red[0][warp] = ce; red[1][warp] = nv; red[2][warp] = A; }
    __syncthreads();
    if (threadIdx.x < 2 * (kFoot + 2)) {               // range tables (cx0 / cy0 are monotone along the tile)
        const int which = threadIdx.x / (kFoot + 2), rr = threadIdx.x % (kFoot + 2);
        const int* src = which ? cy0 : cx0;
        int n = 0;
        for (int i = 0; i < kTile; ++i) n += src[i] < rr ? 1 : 0;
        (which ? fy : fx)[rr] = n;
    }
    if (threadIdx.x == 64 || threadIdx.x == 65) {     // footprint: source cell of the last column / row inside the image, plus its right neighbour
        const int which = threadIdx.x - 64, lim = which ? H - ty0 : W - tx0, last = (lim < kTile ? lim : kTile) - 1;
        const int n = (which ? cy0 : cx0)[last] + 2;
        nf[which] = n < kFoot ? n : kFoot;
    }
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int i = 0; i < kLossThreads / 32; ++i) s += red[threadIdx.x][i];
        atomicAdd(threadIdx.x == 0 ? acc : (threadIdx.x == 1 ? acc + 1 : acc + 2 + b), s);
    }
    if (threadIdx.x == 3 && flags[0]) atomicOr(reinterpret_cast<int*>(acc + 2 + B) + b, 1);
    if (threadIdx.x == 4 && flags[1]) atomicOr(reinterpret_cast<int*>(acc + 2 + 2 * B) + b, 1);
    __syncthreads();
    // ---- x pass: T[row][rx][k] = sum over the output columns whose bilinear footprint contains low-res column fx0 + rx:
    //      columns with x0 == rx (low weight) and columns with x0 == rx - 1 (high weight).
    //      One thread per (output row, low-res column INSIDE the footprint), the 7 classes in registers: at scale 4 a tile touches
    //      ~10 of the kFoot = 18 columns, and one work item per (row, column, class) over all 18 -- 4032 items with an index decode
    //      and two loop set-ups each -- was 39 % of the kernel's instructions (profiles/ncu_r2_seg_loss_source_lines.txt).
    //      Same summation order per element as before: bit-identical.
    const int nfx = nf[0], nfy = nf[1];
    for (int i = threadIdx.x; i < kTile * nfx; i += blockDim.x) {
        const int row = i / nfx, rx = i - row * nfx;
        float s[kNC];
#pragma unroll
        for (int k = 0; k < kNC; ++k) s[k] = 0.f;
        for (int cc = fx[rx]; cc < fx[rx + 1]; ++cc) {
            const float wgt = wx0[cc];
#pragma unroll
            for (int k = 0; k < kNC; ++k) s[k] += wgt * G[row][cc * kNC + k];
        }
        if (rx > 0)
            for (int cc = fx[rx - 1]; cc < fx[rx]; ++cc) {
                const float wgt = wx1[cc];
#pragma unroll
                for (int k = 0; k < kNC; ++k) s[k] += wgt * G[row][cc * kNC + k];
            }
#pragma unroll
        for (int k = 0; k < kNC; ++k) T[row][rx][k] = s[k];
    }
    __syncthreads();
    // ---- y pass, straight into the global low-resolution gradient (cells on tile borders are shared with the neighbours)
    for (int i = threadIdx.x; i < nfy * nfx; i += blockDim.x) {
        const int ry = i / nfx, rx = i - ry * nfx;
        float vsum[kNC];
#pragma unroll
        for (int k = 0; k < kNC; ++k) vsum[k] = 0.f;
        for (int rr = fy[ry]; rr < fy[ry + 1]; ++rr) {
            const float wgt = wy0[rr];
#pragma unroll
            for (int k = 0; k < kNC; ++k) vsum[k] += wgt * T[rr][rx][k];
        }
        if (ry > 0)
            for (int rr = fy[ry - 1]; rr < fy[ry]; ++rr) {
                const float wgt = wy1[rr];
#pragma unroll
                for (int k = 0; k < kNC; ++k) vsum[k] += wgt * T[rr][rx][k];
            }
        const int yy = fy0 + ry, xx = fx0 + rx;
        if (yy < h && xx < w) {
            float* gp = gdir + (((int64_t)b * h + yy) * w + xx) * kNCP;
#pragma unroll
            for (int k = 0; k < kNC; ++k)
                if (vsum[k] != 0.f) atomicAdd(gp + k, vsum[k]);
        }
    }
}

// out[0] = loss, out[1] = gradient scale (factor / n_valid), out[2] = CE, out[3] = n_valid
__global__ void seg_loss_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ aux_scores, float* __restrict__ out, int B) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float ce_sum = acc[0], nvalid = acc[1];
    const int* has_fg = reinterpret_cast<const int*>(acc + 2 + B);
    const int* has_bg = reinterpret_cast<const int*>(acc + 2 + 2 * B);
    float mod = 0.f;
    for (int b = 0; b < B; ++b) {
        float l1 = 0.f;
        for (int c = 0; c < kNC; ++c) {
            const float onehot = (c == 0) ? (has_bg[b] ? 1.f : 0.f) : ((c == 1) ? (has_fg[b] ? 1.f : 0.f) : 0.f);
            l1 += 1.f / (1.f + expf(fabsf(aux_scores[b * kNC + c] - onehot)));
        }
        l1 /= (2.f * B);
        mod += (1.f - l1 / 7.f) * acc[2 + b];
    }
    const float ce = ce_sum / nvalid;                 // NaN when nothing is valid, like F.cross_entropy
    const float factor = mod / (nvalid + (float)B);
    out[0] = ce * factor;
    out[1] = factor / nvalid;
    out[2] = ce;
    out[3] = nvalid;
}

__global__ void seg_loss_bwd_kernel(const float* __restrict__ gdir, const float* __restrict__ fin, const float* __restrict__ upstream,
                                    float* __restrict__ dlogits, int64_t n) {
    const float s = fin[1] * (upstream ? upstream[0] : 1.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dlogits[i] = gdir[i] * s;
}

// ---------------------------------------------------------------------------------------------
// optimiser: global grad-norm clip + SGD(momentum, weight decay) over ONE flat fp32 parameter buffer
// ---------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ g, int64_t n, float gscale, double* __restrict__ out) {
    __shared__ float red[32];
    float s = 0.f;
    const int64_t n4 = n / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) for (int64_t i = n4 * 4; i < n; ++i) s += g[i] * g[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) atomicAdd(out, (double)t * (double)gscale * (double)gscale);
    }
}

// g' = g*gscale*clip + wd*p ; m = mu*m + g' (m starts at 0, i.e. m = g' on the first step like torch.optim.SGD) ;
// p -= lr*m ; optionally g = 0 and bf16 shadow copy.  lr is read from device memory so that a captured CUDA graph
// of the step follows the poly schedule without re-capture.
__global__ void sgd_step_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, int64_t n,
                                const double* __restrict__ sumsq, float gscale, float max_norm, const float* __restrict__ lr_dev,
                                float mu, float wd, int zero_grad, __nv_bfloat16* __restrict__ shadow) {
    const float lr = *lr_dev;
    const float total = (float)sqrt(*sumsq);
    float clip = max_norm > 0.f ? max_norm / (total + 1e-6f) : 1.f;
    clip = fminf(clip, 1.f) * gscale;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float pv = p[i];
        const float gv = g[i] * clip + wd * pv;
        const float mv = mu * m[i] + gv;
        const float np = pv - lr * mv;
        m[i] = mv;
        p[i] = np;
        if (zero_grad) g[i] = 0.f;
        if (shadow) shadow[i] = __float2bfloat16_rn(np);
    }
}

// Channels-last bf16 copies of the k x k (k > 1) convolution weights, refreshed once per step right after the update:
// the library convolutions want (Cout, kh, kw, Cin) operands for NHWC activations, and converting the (Cout, Cin, kh, kw)
// shadow per call cost one small permute kernel in front of every 3x3 convolution (~170 per step, each on a latency-bound
// chain).  One "row" = one output channel of one weight: Cin*kk contiguous floats in, Cin*kk contiguous bf16 out, permuted
// through shared memory so both sides are coalesced.  table[e] = {src offset (floats), dst offset (elements), Cin, kk};
// row_start[e] = first global row of entry e (row_start[n_entries] = total rows).
// One WARP per row (round 2: one block per row with two block barriers -- 256 threads for rows of 288 floats -- took 119 us for
// 108 MB in + 54 MB out): 16-byte loads into the warp's slab, 16-byte stores of 8 consecutive input channels of one tap; the lanes
// walk the output vectors tap-fastest so that the stride-kk shared-memory reads of a warp fall into distinct banks.
constexpr int kClWarps = 8;
__global__ void __launch_bounds__(kClWarps * 32)
shadow_cl_kernel(const float* __restrict__ p, __nv_bfloat16* __restrict__ out, const int64_t* __restrict__ table,
                 const int64_t* __restrict__ row_start, int n_entries, int slab /*floats per warp, multiple of 4*/) {
    extern __shared__ __align__(16) float row_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sm = row_sm + (size_t)warp * slab;
    const int64_t total = row_start[n_entries];
    for (int64_t row = (int64_t)blockIdx.x * kClWarps + warp; row < total; row += (int64_t)gridDim.x * kClWarps) {
        int lo = 0, hi = n_entries - 1;                     // last entry with row_start <= row
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (row_start[mid] <= row) lo = mid; else hi = mid - 1;
        }
        const int64_t* e = table + (int64_t)lo * 4;
        const int cin = (int)e[2], kk = (int)e[3], L = cin * kk;
        const int64_t local = row - row_start[lo];
        const float* src = p + e[0] + local * L;
        __nv_bfloat16* dst = out + e[1] + local * L;
        if ((L & 3) == 0 && ((uintptr_t)src & 15) == 0) {
            for (int j = lane; j < (L >> 2); j += 32)
                reinterpret_cast<float4*>(sm)[j] = __ldg(reinterpret_cast<const float4*>(src) + j);
        } else {
            for (int j = lane; j < L; j += 32) sm[j] = src[j];
        }
        __syncwarp();
        if ((cin & 7) == 0 && ((uintptr_t)dst & 15) == 0) {
            const int cchunks = cin >> 3;
            for (int v = lane; v < cchunks * kk; v += 32) {
                const int cc = v / kk, t = v - cc * kk;        // output vector: tap t, input channels cc*8 .. cc*8+7
                const float* s8 = sm + (cc * 8) * kk + t;
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(s8[(2 * q) * kk], s8[(2 * q + 1) * kk]);
                    w[q] = *reinterpret_cast<uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(dst + t * cin + cc * 8) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
            for (int j = lane; j < L; j += 32) {
                const int t = j / cin, ci = j - t * cin;
                dst[j] = __float2bfloat16_rn(sm[ci * kk + t]);
            }
        }
        __syncwarp();
    }
}

// dst_e[i] += float(src_e[i]) for a LIST of tensors in one launch: the fp32 accumulation of the library's bf16 weight gradients into
// the flat gradient buffer (one torch elementwise launch per convolution before: 218 launches per step).
// table[e] = {src pointer, dst pointer, numel, Cin, kk}: kk = 0: same element order; kk > 0: src is the (Cout,kh,kw,Cin) channels-last
// layout the library returns for k x k weights, dst the (Cout,Cin,kh,kw) parameter order.
// chunk_start[e] = first chunk of entry e (chunk_start[n] = total chunks); a chunk is 4096 elements, or -- kk > 0 and a row of
// Cin*kk elements fits -- max(1, 4096 / (Cin*kk)) whole rows (rss_accum_chunks() is the one definition, host and device).
// Round 2 walked the k x k entries in destination order and gathered the bf16 source element by element (stride Cin, two integer
// divisions each): 144 us for 54 MB in + 216 MB read-modify-write.  Whole rows are contiguous on BOTH sides, so a chunk of rows is
// staged through shared memory: 16-byte source loads, 16-byte read-modify-writes.
constexpr int kAccChunk = 4096;
__host__ __device__ inline int64_t accum_rows_per_chunk(int64_t cin, int64_t kk) {
    const int64_t L = cin * kk;
    return (kk > 0 && L > 0 && L <= kAccChunk) ? (kAccChunk / L > 1 ? kAccChunk / L : 1) : 0;      // 0: element chunks
}
__global__ void __launch_bounds__(256)
accum_list_kernel(const int64_t* __restrict__ table, const int64_t* __restrict__ chunk_start, int n_entries) {
    __shared__ __align__(16) float stage[kAccChunk];
    const int64_t total = chunk_start[n_entries];
    for (int64_t chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
        int lo = 0, hi = n_entries - 1;                     // last entry with chunk_start <= chunk
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (chunk_start[mid] <= chunk) lo = mid; else hi = mid - 1;
        }
        const int64_t* e = table + (int64_t)lo * 5;
        const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(e[0]);
        float* dst = reinterpret_cast<float*>(e[1]);
        const int64_t n = e[2], cl = chunk - chunk_start[lo];
        const int cin = (int)e[3], kk = (int)e[4];
        const int64_t rpc = accum_rows_per_chunk(cin, kk);
        if (kk == 0) {
            const int64_t base = cl * kAccChunk;
            for (int64_t i = base + threadIdx.x; i < n && i < base + kAccChunk; i += blockDim.x) dst[i] += __bfloat162float(src[i]);
        } else if (rpc == 0) {                              // rows longer than the staging buffer: element-wise gather
            const int64_t base = cl * kAccChunk;
            const int row = cin * kk;
            for (int64_t i = base + threadIdx.x; i < n && i < base + kAccChunk; i += blockDim.x) {
                const int co = (int)(i / row), rem = (int)(i - (int64_t)co * row), ci = rem / kk, t = rem - ci * kk;
                dst[i] += __bfloat162float(src[((int64_t)co * kk + t) * cin + ci]);
            }
        } else {
            const int L = cin * kk;
            const int64_t rows = n / L, r0 = cl * rpc;
            const int nr = (int)(rows - r0 < rpc ? rows - r0 : rpc), E = nr * L;
            const __nv_bfloat16* s = src + r0 * L;
            float* d = dst + r0 * L;
            __syncthreads();                                // the previous chunk's readers of `stage` are done
            if ((E & 7) == 0 && ((uintptr_t)s & 15) == 0) {
                for (int j = threadIdx.x; j < (E >> 3); j += blockDim.x) {
                    float v[8];
                    load8(s + j * 8, v);
                    reinterpret_cast<float4*>(stage)[2 * j] = make_float4(v[0], v[1], v[2], v[3]);
                    reinterpret_cast<float4*>(stage)[2 * j + 1] = make_float4(v[4], v[5], v[6], v[7]);
                }
            } else {
                for (int j = threadIdx.x; j < E; j += blockDim.x) stage[j] = __bfloat162float(s[j]);
            }
            __syncthreads();
            auto permuted = [&](int j) {                    // destination element j (row, ci, t) -> its staged (row, t, ci) value
                const int rr = j / L, rem = j - rr * L, ci = rem / kk, t = rem - ci * kk;
                return stage[rr * L + t * cin + ci];
            };
            if ((E & 3) == 0 && ((uintptr_t)d & 15) == 0) {
                for (int j = threadIdx.x; j < (E >> 2); j += blockDim.x) {
                    float4 g = reinterpret_cast<float4*>(d)[j];
                    g.x += permuted(4 * j); g.y += permuted(4 * j + 1); g.z += permuted(4 * j + 2); g.w += permuted(4 * j + 3);
                    reinterpret_cast<float4*>(d)[j] = g;
                }
            } else {
                for (int j = threadIdx.x; j < E; j += blockDim.x) d[j] += permuted(j);
            }
        }
    }
}

// number of accum_list_kernel chunks of one table entry
extern "C" int64_t rss_accum_chunks(int64_t numel, int64_t cin, int64_t kk) {
    const int64_t rpc = accum_rows_per_chunk(cin, kk);
    if (rpc == 0) return (numel + kAccChunk - 1) / kAccChunk;
    const int64_t rows = numel / (cin * kk);
    return (rows + rpc - 1) / rpc;
}

// transposed copies for the data-gradient operand of the fused conv: out[(ci*kk + t)*Cout + co] = w[(co*Cin + ci)*kk + t]
__global__ void shadow_t_kernel(const float* __restrict__ p, __nv_bfloat16* __restrict__ out, const int64_t* __restrict__ table) {
    const int64_t* e = table + (int64_t)blockIdx.y * 5;
    const int cout = (int)e[2], cin = (int)e[3], kk = (int)e[4], n = cout * cin * kk;
    const float* src = p + e[0];
    __nv_bfloat16* dst = out + e[1];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int co = j % cout, r = j / cout, t = r % kk, ci = r / kk;
        dst[j] = __float2bfloat16_rn(src[((int64_t)co * cin + ci) * kk + t]);
    }
}

}  // namespace rss

using namespace rss;

extern "C" int rss_accum_bf16_list(const int64_t* table, const int64_t* chunk_start, int n_entries, int64_t total_chunks, cudaStream_t st) {
    if (n_entries <= 0 || total_chunks <= 0) return RSS_ERR_SHAPE;
    int grid = (int)(total_chunks < (int64_t)num_sms() * 8 ? total_chunks : (int64_t)num_sms() * 8);
    accum_list_kernel<<<grid, 256, 0, st>>>(table, chunk_start, n_entries);
    return check_launch();
}

extern "C" int rss_shadow_t_refresh(const float* params, void* shadow_t, const int64_t* table, int n_entries, cudaStream_t st) {
    if (n_entries <= 0 || n_entries > 65535) return RSS_ERR_SHAPE;
    shadow_t_kernel<<<dim3(8, n_entries), 256, 0, st>>>(params, (__nv_bfloat16*)shadow_t, table);
    return check_launch();
}

extern "C" int rss_shadow_cl_refresh(const float* params, void* shadow_cl, const int64_t* table, const int64_t* row_start,
                                     int n_entries, int max_row_floats, cudaStream_t st) {
    if (n_entries <= 0 || max_row_floats <= 0 || max_row_floats > 6 * 1024) return RSS_ERR_SHAPE;
    const int slab = (max_row_floats + 3) & ~3;
    const size_t smem = (size_t)kClWarps * slab * sizeof(float);
    static size_t attr_smem[16] = {0};                      // per device: opt-in dynamic shared memory already granted
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && (dev < 0 || dev >= 16 || attr_smem[dev] < smem)) {
        cudaError_t e = cudaFuncSetAttribute(shadow_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { g_last_cuda_error = (int)e; return RSS_ERR_CUDA; }
        if (dev >= 0 && dev < 16) attr_smem[dev] = smem;
    }
    shadow_cl_kernel<<<num_sms() * 3, kClWarps * 32, smem, st>>>(params, (__nv_bfloat16*)shadow_cl, table, row_start, n_entries, slab);
    return check_launch();
}

extern "C" size_t rss_seg_loss_acc_floats(int B) { return (size_t)(2 + 3 * B); }

extern "C" int rss_seg_loss_fwd(const float* logits_lr, const int64_t* labels, const float* aux_scores, float* acc_ws,
                                float* gdir, float* out4, int B, int h, int w, int scale, int ignore_index, cudaStream_t st) {
    if (B <= 0 || h <= 0 || w <= 0 || scale < 2) return RSS_ERR_SHAPE;
    const int H = h * scale, W = w * scale;
    cudaError_t e = cudaMemsetAsync(acc_ws, 0, rss_seg_loss_acc_floats(B) * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(gdir, 0, (size_t)B * h * w * kNCP * sizeof(float), st);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return RSS_ERR_CUDA; }
    const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    dim3 grid(((H + kTile - 1) / kTile) * ((W + kTile - 1) / kTile), B);
    seg_loss_fwd_kernel<<<grid, kLossThreads, 0, st>>>(logits_lr, labels, acc_ws, gdir, B, h, w, H, W, sy, sx, ignore_index);
    seg_loss_finalize_kernel<<<1, 32, 0, st>>>(acc_ws, aux_scores, out4, B);
    return check_launch();
}

extern "C" int rss_seg_loss_bwd(const float* gdir, const float* out4, const float* upstream, float* dlogits_lr,
                                int B, int h, int w, cudaStream_t st) {
    const int64_t n = (int64_t)B * h * w * kNCP;
    if (n <= 0) return RSS_ERR_SHAPE;
    int grid = (int)((n + 255) / 256);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    seg_loss_bwd_kernel<<<grid, 256, 0, st>>>(gdir, out4, upstream, dlogits_lr, n);
    return check_launch();
}

extern "C" int rss_grad_sumsq(const float* grads, int64_t n, float grad_scale, double* sumsq, cudaStream_t st) {
    if (n <= 0) return RSS_ERR_SHAPE;
    cudaError_t e = cudaMemsetAsync(sumsq, 0, sizeof(double), st);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return RSS_ERR_CUDA; }
    sumsq_kernel<<<num_sms() * 4, 256, 0, st>>>(grads, n, grad_scale, sumsq);
    return check_launch();
}

extern "C" int rss_sgd_step(float* params, float* grads, float* momentum_buf, int64_t n, const double* sumsq, float grad_scale,
                            float max_norm, const float* lr, float momentum, float weight_decay, int zero_grad,
                            void* bf16_shadow, cudaStream_t st) {
    if (n <= 0 || !lr) return RSS_ERR_SHAPE;
    sgd_step_kernel<<<num_sms() * 8, 256, 0, st>>>(params, grads, momentum_buf, n, sumsq, grad_scale, max_norm, lr, momentum,
                                                   weight_decay, zero_grad, (__nv_bfloat16*)bf16_shadow);
    return check_launch();
}
