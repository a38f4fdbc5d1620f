// Permutohedral-lattice bilateral filter on the device: the CUDA replacement of the reference's only native
// component, the SWIG-wrapped C++ filter behind DenseEnergyLoss (SCD-AAAI2023/utils/losses.py:52-91), which today costs a
// GPU -> CPU -> GPU round trip per training step:
//   SCD-AAAI2023/wrapper/bilateralfilter/bilateralfilter.hpp:12      void bilateralfilter_batch(images, ins, outs, N, K, H, W, sigmargb, sigmaxy)
//   SCD-AAAI2023/wrapper/bilateralfilter/bilateralfilter.cpp:4-55    features (x, y, r, g, b) / sigma, one lattice per image, one class plane at a time
//   SCD-AAAI2023/wrapper/bilateralfilter/permutohedral.cpp:116-300   Permutohedral::init (embedding, hash table of lattice points, blur neighbours)
//   SCD-AAAI2023/wrapper/bilateralfilter/permutohedral.cpp:490-553   Permutohedral::compute (splat, blur along d+1 directions, slice)
//
// RESULTS ARE BIT-IDENTICAL TO THE COMPILED REFERENCE (tests/test_gpu_bilateral.py), which takes three things:
//   1. every float multiply / add / divide is a separately rounded IEEE operation in the reference's operand order (__fmul_rn /
//      __fadd_rn / __fsub_rn / __fdiv_rn: nothing here may be contracted into an FMA);
//   2. the splat adds the pixels of a lattice point in RASTER ORDER (float addition is not associative), so there are no float
//      atomics: the (lattice point, pixel) pairs are brought into (point, raster) order by a stable LSD radix sort and one warp
//      walks the list of one lattice point, lanes = class planes;
//   3. the reference's SSE build embeds pixels four at a time and also CREATES the lattice points of the 1-3 zero-feature padding
//      lanes when H*W is not a multiple of 4; they carry no signal but relay the blur, so they are created here too.
// The identifiers of the lattice points differ from the reference's (order of hash insertion); no result depends on them.
//
// B200 notes.  One hash table for the whole batch (the image index is part of the key), keys are 16 bytes (5 x int16 lattice
// coordinates + image) and are claimed with ONE 128-bit compare-and-swap (ATOMG.E.CAS.128, sm_90+): no lock word, no spinning, and
// a thread first checks the slot with a plain 16-byte load because neighbouring pixels share their lattice points (the CAS is the
// exception, not the rule).  All K class planes go through the lattice together (the reference makes K passes with value_size 1;
// the planes are independent, so the arithmetic per plane is the same).  Everything is stream-ordered and sized by upper bounds
// from the host: the number of lattice points M stays on the device (grid-stride loops read it), there is no host synchronisation.
// HBM-bound by design: algorithmic bytes = images + ins + outs = (3 + 2K) * 4 * N*H*W; the lattice traffic (hash table, sort,
// values) lives in the 126 MB L2 at the reference's working sizes (N=2, K=21, 160x160: 100 MB of workspace, ~12 MB touched).
#include <math.h>
#include "common.cuh"

namespace rss {

constexpr int PD = 5;            // feature dimensions
constexpr int PV = PD + 1;       // simplex vertices
constexpr int SORT_TILE = 2048;  // entries per block of the radix-sort kernels (256 threads x 8)

struct alignas(16) LKey { unsigned w0, w1, w2, w3; };   // coordinates 0..4 as uint16 pairs, w3 = image index; empty slot = all ones

struct EmbedConsts { float scale[PD]; float inv_v, v, alpha; };

__device__ __forceinline__ bool key_eq(const LKey& a, const LKey& b) { return a.w0 == b.w0 && a.w1 == b.w1 && a.w2 == b.w2 && a.w3 == b.w3; }
__device__ __forceinline__ bool key_empty(const LKey& a) { return a.w3 == 0xffffffffu; }

__device__ __forceinline__ unsigned key_hash(const LKey& k) {
    unsigned long long h = ((unsigned long long)k.w1 << 32 | k.w0) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 32;
    h += ((unsigned long long)k.w3 << 32 | k.w2);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
    return (unsigned)h;
}

__device__ __forceinline__ LKey load_key(const LKey* p) {
    // one 16-byte access, cached in L1: a slot only ever goes EMPTY -> key, so a stale EMPTY just sends the thread to the CAS (which is
    // authoritative) and a cached key is final.  L2-only loads (first version) hot-spotted the few L2 lines that every warp working
    // on the same image region probes.
    const uint4 v = __ldca(reinterpret_cast<const uint4*>(p));
    return LKey{v.x, v.y, v.z, v.w};
}

// ---- 1. embedding + creation of the lattice points ---------------------------------------------------------------------------
// one thread per (image, pixel slot); slots P..P4-1 of an image are the zero-feature padding lanes (created, not recorded)
__global__ void pl_embed_kernel(const float* __restrict__ images, int N, int P, int P4, int W, float sigmargb, float sigmaxy,
                                EmbedConsts ec, LKey* table, unsigned mask, int* __restrict__ id_of_slot, LKey* __restrict__ key_of_id,
                                int* counters, int* __restrict__ vertex, float* __restrict__ weight) {
    const long long total = (long long)N * P4;
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count: the lanes of a warp de-duplicate their keys with __match_any_sync before anyone probes the table
    for (long long base0 = (long long)blockIdx.x * blockDim.x; base0 < total; base0 += (long long)gridDim.x * blockDim.x) {
        const long long idx = base0 + threadIdx.x;
        const bool active = idx < total;
        // ids are handed out per block: one global atomic per 256 pixels (one per created point serialised the kernel on a single L2
        // address: 65 us of its 81 us at N=8, 224x224)
        __shared__ int blk_created, blk_base;
        if (threadIdx.x == 0) blk_created = 0;
        __syncthreads();
        int my_slot[PV], my_rank[PV];
        const int b = active ? (int)(idx / P4) : 0, p = active ? (int)(idx - (long long)b * P4) : P4;
        float f[PD] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (p < P) {
            const int y = p / W, x = p - y * W;
            const float* im = images + (size_t)b * 3 * P + p;
            f[0] = __fdiv_rn((float)x, sigmaxy);
            f[1] = __fdiv_rn((float)y, sigmaxy);
            f[2] = __fdiv_rn(__ldg(im), sigmargb);
            f[3] = __fdiv_rn(__ldg(im + P), sigmargb);
            f[4] = __fdiv_rn(__ldg(im + 2 * (size_t)P), sigmargb);
        }
        // elevate (permutohedral.cpp:187-193)
        float el[PV], base[PV];
        float run = 0.f;
#pragma unroll
        for (int j = PD; j > 0; --j) {
            const float cf = __fmul_rn(f[j - 1], ec.scale[j - 1]);
            el[j] = __fsub_rn(run, __fmul_rn((float)j, cf));
            run = __fadd_rn(run, cf);
        }
        el[0] = run;
        // nearest remainder-0 point, round half to even (:196-206)
        float coord_sum = 0.f;
#pragma unroll
        for (int i = 0; i < PV; ++i) {
            const float q = rintf(__fmul_rn(ec.inv_v, el[i]));
            base[i] = __fmul_rn(q, ec.v);
            coord_sum = __fadd_rn(coord_sum, q);
        }
        // ranks of the residuals (:209-219); small exact integers, kept as int
        float res[PV];
        int rank[PV];
#pragma unroll
        for (int i = 0; i < PV; ++i) { res[i] = __fsub_rn(el[i], base[i]); rank[i] = 0; }
#pragma unroll
        for (int i = 0; i < PD; ++i)
#pragma unroll
            for (int j = i + 1; j < PV; ++j) {
                const int lt = res[i] < res[j] ? 1 : 0;
                rank[i] += lt;
                rank[j] += 1 - lt;
            }
        // off-plane correction (:222-228)
        const int isum = (int)coord_sum;
        int ibase[PV];
#pragma unroll
        for (int i = 0; i < PV; ++i) {
            rank[i] += isum;
            float adj = 0.f;
            if (rank[i] < 0) { rank[i] += PV; adj = ec.v; }
            else if (rank[i] >= PV) { rank[i] -= PV; adj = -ec.v; }
            base[i] = __fadd_rn(base[i], adj);
            rank[i] = min(max(rank[i], 0), PD);            // (the reference indexes out of bounds beyond this; never reached for finite input)
            ibase[i] = (int)base[i];
        }
        // barycentric weights (:231-247).  The reference scatters +t / -t into b[d-rank], b[d-rank+1] in coordinate order; the ranks
        // are a permutation, every b[p] receives exactly one +t and one -t, and (0 + a) - c == (0 - c) + a in IEEE arithmetic, so
        // b[p] = t(rank d-p) - t(rank d-p+1), b[0] = t(rank d) + (1 - t(rank 0)) bit for bit.
        float t_of_rank[PV];
#pragma unroll
        for (int r = 0; r < PV; ++r) t_of_rank[r] = 0.f;
#pragma unroll
        for (int i = 0; i < PV; ++i) {
            const float t = __fmul_rn(__fsub_rn(el[i], base[i]), ec.inv_v);
#pragma unroll
            for (int r = 0; r < PV; ++r)
                if (rank[i] == r) t_of_rank[r] = t;
        }
        float bary[PV];
        bary[0] = __fadd_rn(t_of_rank[PD], __fadd_rn(1.0f, __fsub_rn(0.f, t_of_rank[0])));
#pragma unroll
        for (int q = 1; q < PV; ++q) bary[q] = __fsub_rn(t_of_rank[PD - q], t_of_rank[PD - q + 1]);
        // the simplex vertices (:254-262): vertex r = base + row r of the canonical simplex, indexed by rank
#pragma unroll
        for (int r = 0; r < PV; ++r) {
            unsigned short c[PD];
#pragma unroll
            for (int i = 0; i < PD; ++i) c[i] = (unsigned short)(short)(ibase[i] + (rank[i] <= PD - r ? r : r - PV));
            // inactive lanes carry a per-lane key no pixel can have (image index 0xfffffffe) and never touch the table
            const LKey key = active ? LKey{(unsigned)c[0] | (unsigned)c[1] << 16, (unsigned)c[2] | (unsigned)c[3] << 16, (unsigned)c[4], (unsigned)b}
                                    : LKey{(unsigned)lane, 0u, 0u, 0xfffffffeu};
            const LKey empty{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
            // neighbouring pixels share most of their lattice points: only the first lane of every group of equal keys probes
            const unsigned grp = __match_any_sync(0xffffffffu, (unsigned long long)key.w1 << 32 | key.w0) &
                                 __match_any_sync(0xffffffffu, (unsigned long long)key.w3 << 32 | key.w2);
            const int leader = __ffs(grp) - 1;
            unsigned h = key_hash(key) & mask;
            my_rank[r] = -1;
            if (active && lane == leader) {
                for (;;) {
                    LKey cur = load_key(table + h);
                    if (key_empty(cur)) cur = atomicCAS(table + h, empty, key);
                    if (key_empty(cur)) {                  // this thread created the lattice point
                        my_rank[r] = atomicAdd(&blk_created, 1);
                        my_slot[r] = (int)h;
                        break;
                    }
                    if (key_eq(cur, key)) break;
                    h = (h + 1) & mask;
                }
            }
            h = __shfl_sync(0xffffffffu, h, leader);
            if (p < P) {
                const size_t e = ((size_t)b * P + p) * PV + r;
                vertex[e] = (int)h;                        // hash slot for now; pl_entries_kernel turns it into the point id
                weight[e] = bary[r];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) blk_base = blk_created ? atomicAdd(counters, blk_created) : 0;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < PV; ++r)
            if (my_rank[r] >= 0) {
                unsigned short c[PD];
#pragma unroll
                for (int i = 0; i < PD; ++i) c[i] = (unsigned short)(short)(ibase[i] + (rank[i] <= PD - r ? r : r - PV));
                const int id = blk_base + my_rank[r];
                id_of_slot[my_slot[r]] = id;
                key_of_id[id] = LKey{(unsigned)c[0] | (unsigned)c[1] << 16, (unsigned)c[2] | (unsigned)c[3] << 16, (unsigned)c[4], (unsigned)b};
            }
    }
}

// ---- 2. (point id, entry) pairs for the sort ---------------------------------------------------------------------------------
// one block per sort tile; also leaves the tile's histogram of the first digit (saves the first pl_sort_hist_kernel launch)
__global__ void pl_entries_kernel(int* __restrict__ vertex, const int* __restrict__ id_of_slot, unsigned* __restrict__ keys,
                                  unsigned* __restrict__ vals, unsigned E, unsigned* __restrict__ hist, int nblk) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned t0 = blockIdx.x * 2048u;
    for (int i = threadIdx.x; i < 2048; i += 256) {
        const unsigned e = t0 + i;
        if (e < E) {
            const int id = id_of_slot[vertex[e]];
            vertex[e] = id;
            keys[e] = (unsigned)id;
            vals[e] = e;
            atomicAdd(&h[(unsigned)id & 255u], 1u);
        }
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// ---- 3. stable LSD radix sort of the pairs by point id, 8 bits per pass --------------------------------------------------------
__global__ void pl_sort_hist_kernel(const unsigned* __restrict__ keys, unsigned n, int shift, unsigned* __restrict__ hist, int nblk) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned t0 = blockIdx.x * SORT_TILE;
    for (int i = threadIdx.x; i < SORT_TILE; i += 256)
        if (t0 + i < n) atomicAdd(&h[(keys[t0 + i] >> shift) & 255u], 1u);
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of the digit histograms (256 x #tiles entries), two levels: every block scans a chunk of 4096 in place and leaves its
// total in sums[chunk]; the last block to finish scans the totals; the scatter kernel adds sums[i / 4096] when it reads entry i.
// (First version: ONE block over the whole array = 120 us of the 335 us call at N=2, 160x160 and 11 ms of 18 ms at N=16, 512x512.)
constexpr int SCAN_CHUNK = 4096;
// exclusive scan in place by ONE block of 1024 threads (the chunk totals: a few hundred entries)
__device__ __forceinline__ void block_scan_inplace(unsigned* a, unsigned len, unsigned* wsum, unsigned* carry) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) *carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < len; base += 1024) {
        const unsigned i = base + threadIdx.x;
        const unsigned v = i < len ? __ldcg(a + i) : 0u;
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned s = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += t;
            }
            wsum[lane] = s;                                 // inclusive over warps
        }
        __syncthreads();
        const unsigned before = *carry + (warp > 0 ? wsum[warp - 1] : 0u);
        if (i < len) a[i] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) *carry += wsum[31];
        __syncthreads();
    }
}

__global__ void pl_scan_chunks_kernel(unsigned* __restrict__ a, unsigned len, unsigned* sums, int* ticket) {
    __shared__ unsigned wsum[32];
    __shared__ unsigned carry;
    __shared__ int last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned i0 = blockIdx.x * SCAN_CHUNK + threadIdx.x * 4;
    unsigned v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = i0 + q < len ? a[i0 + q] : 0u;
    const unsigned tsum = v[0] + v[1] + v[2] + v[3];
    unsigned inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned sv = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, sv, o);
            if (lane >= o) sv += t;
        }
        wsum[lane] = sv;
    }
    __syncthreads();
    unsigned run = inc - tsum + (warp > 0 ? wsum[warp - 1] : 0u);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (i0 + q < len) a[i0 + q] = run;
        run += v[q];
    }
    // the block that finishes last scans the chunk totals (one launch per pass instead of two)
    if (threadIdx.x == 0) {
        sums[blockIdx.x] = wsum[31];
        __threadfence();
        last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        block_scan_inplace(sums, gridDim.x, wsum, &carry);
    }
}

// A block scatters its tile in index order: warp w owns 256 consecutive entries and takes them 32 at a time; inside a round the
// rank among equal digits comes from __match_any_sync, across rounds and warps from per-warp digit counters -> the pass is stable.
__global__ void pl_sort_scatter_kernel(const unsigned* __restrict__ keys, const unsigned* __restrict__ vals, unsigned* __restrict__ okeys,
                                       unsigned* __restrict__ ovals, unsigned n, int shift, const unsigned* __restrict__ hist,
                                       const unsigned* __restrict__ sums, int nblk) {
    __shared__ unsigned wcnt[8][256];
    __shared__ unsigned gbase[256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    unsigned k[8], v[8], rk[8];
    const unsigned t0 = blockIdx.x * SORT_TILE + w * 256;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const unsigned idx = t0 + r * 32 + lane;
        const bool valid = idx < n;
        k[r] = valid ? keys[idx] : 0u;
        v[r] = valid ? vals[idx] : 0u;
        const unsigned d = valid ? (k[r] >> shift) & 255u : 256u;
        const unsigned m = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(m) - 1;
        unsigned base = 0;
        if (valid && lane == leader) {
            base = wcnt[w][d];
            wcnt[w][d] = base + __popc(m);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        rk[r] = base + __popc(m & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {
        const int d = threadIdx.x;
        unsigned run = 0;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) {
            const unsigned c = wcnt[ww][d];
            wcnt[ww][d] = run;
            run += c;
        }
        const unsigned hi = (unsigned)d * (unsigned)nblk + blockIdx.x;
        gbase[d] = hist[hi] + sums[hi / SCAN_CHUNK];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const unsigned idx = t0 + r * 32 + lane;
        if (idx < n) {
            const unsigned d = (k[r] >> shift) & 255u;
            const unsigned pos = gbase[d] + wcnt[w][d] + rk[r];
            okeys[pos] = k[r];
            ovals[pos] = v[r];
        }
    }
}

// ---- 4. list bounds of every lattice point in the sorted pairs (points made only by padding lanes keep the empty list 0..0) ----
__global__ void pl_segments_kernel(const unsigned* __restrict__ skeys, const unsigned* __restrict__ svals, const float* __restrict__ weight,
                                   unsigned E, int* __restrict__ seg_start, int* __restrict__ seg_end, unsigned* __restrict__ spix,
                                   float* __restrict__ sweight) {
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < E; t += gridDim.x * blockDim.x) {
        const unsigned id = skeys[t];
        const unsigned e = svals[t];
        spix[t] = e / PV;                                   // batch-wide pixel index (the splat reads pixel-major values)
        sweight[t] = weight[e];
        if (t == 0 || skeys[t - 1] != id) seg_start[id] = (int)t;
        if (t == E - 1 || skeys[t + 1] != id) seg_end[id] = (int)(t + 1);
    }
}

// ---- 5. blur neighbours (permutohedral.cpp:279-299) ---------------------------------------------------------------------------
__device__ __forceinline__ int pl_lookup(const LKey* table, unsigned mask, const int* id_of_slot, const LKey& key) {
    unsigned h = key_hash(key) & mask;
    for (;;) {
        const LKey cur = load_key(table + h);
        if (key_empty(cur)) return -1;
        if (key_eq(cur, key)) return id_of_slot[h];
        h = (h + 1) & mask;
    }
}

__global__ void pl_neighbors_kernel(const LKey* __restrict__ table, unsigned mask, const int* __restrict__ id_of_slot,
                                    const LKey* __restrict__ key_of_id, const int* __restrict__ counters, int2* __restrict__ nb, int Mmax) {
    const int M = counters[0];
    const long long total = (long long)M * PV;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int dir = (int)(i / M), id = (int)(i - (long long)dir * M);
        const LKey key = key_of_id[id];
        short c[PD] = {(short)(key.w0 & 0xffffu), (short)(key.w0 >> 16), (short)(key.w1 & 0xffffu), (short)(key.w1 >> 16), (short)(key.w2 & 0xffffu)};
        unsigned short lo[PD], hi[PD];
#pragma unroll
        for (int q = 0; q < PD; ++q) {
            lo[q] = (unsigned short)(short)(c[q] - 1);
            hi[q] = (unsigned short)(short)(c[q] + 1);
            if (q == dir) { lo[q] = (unsigned short)(short)(c[q] + PD); hi[q] = (unsigned short)(short)(c[q] - PD); }
        }
        const LKey k1{(unsigned)lo[0] | (unsigned)lo[1] << 16, (unsigned)lo[2] | (unsigned)lo[3] << 16, (unsigned)lo[4], key.w3};
        const LKey k2{(unsigned)hi[0] | (unsigned)hi[1] << 16, (unsigned)hi[2] | (unsigned)hi[3] << 16, (unsigned)hi[4], key.w3};
        nb[(size_t)dir * Mmax + id] = make_int2(pl_lookup(table, mask, id_of_slot, k1), pl_lookup(table, mask, id_of_slot, k2));
    }
}

// ---- 6. splat (:505-513) -------------------------------------------------------------------------------------------------------
// (N,K,P) -> (N,P,K): a lattice point's list names pixels, and with plane-major values every (entry, plane) load was its own 32-byte
// sector (E*K sectors: 1.6 GB of L2->SM traffic at N=8, 224x224, 11.6 GB from DRAM at N=16, 512x512); pixel-major it is K
// contiguous floats per entry.
constexpr int TP = 128, TK = 32;
__global__ void pl_pixel_major_kernel(const float* __restrict__ ins, float* __restrict__ insT, int K, int P) {
    __shared__ float tile[TK][TP + 1];
    const int b = blockIdx.y, p0 = blockIdx.x * TP, k0 = blockIdx.z * TK;
    const int np = min(TP, P - p0), nk = min(TK, K - k0);
    for (int i = threadIdx.x; i < nk * TP; i += blockDim.x) {
        const int k = i / TP, px = i - k * TP;
        if (px < np) tile[k][px] = __ldg(ins + ((size_t)b * K + k0 + k) * P + p0 + px);
    }
    __syncthreads();
    float* dst = insT + ((size_t)b * P + p0) * K + k0;
    for (int i = threadIdx.x; i < np * nk; i += blockDim.x) {
        const int px = i / nk, k = i - px * nk;
        dst[(size_t)px * K + k] = tile[k][px];
    }
}

// one thread per (lattice point, class plane), planes fastest: the threads of a point read the same list entry (broadcast) and K
// contiguous values; the chain of adds per (point, plane) runs in raster order.  values[(id + 1) * K + k]; row 0 stands for "no
// such lattice point" and stays zero in both buffers.
__device__ __forceinline__ void splat_phase(const float* __restrict__ insT, const unsigned* __restrict__ spix, const float* __restrict__ sweight,
                                            const int* __restrict__ seg_start, const int* __restrict__ seg_end, int M,
                                            float* __restrict__ va, float* __restrict__ vb, int K) {
    const long long total = (long long)M * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int id = (int)(i / K), k = (int)(i - (long long)id * K);
        const int s = seg_start[id], t_end = seg_end[id];
        float acc = 0.f;
        int t = s;
        for (; t + 4 <= t_end; t += 4) {                     // four independent loads in flight, adds in order
            float x[4], w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                w[q] = __ldg(sweight + t + q);
                x[q] = __ldg(insT + (size_t)__ldg(spix + t + q) * K + k);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) acc = __fadd_rn(acc, __fmul_rn(w[q], x[q]));
        }
        for (; t < t_end; ++t) acc = __fadd_rn(acc, __fmul_rn(__ldg(sweight + t), __ldg(insT + (size_t)__ldg(spix + t) * K + k)));
        va[(size_t)(id + 1) * K + k] = acc;
        if (id == 0) { va[k] = 0.f; vb[k] = 0.f; }
    }
}

// ---- 7. blur along one lattice direction, Jacobi style (:515-531) -------------------------------------------------------------
__device__ __forceinline__ void blur_phase(const float* __restrict__ cur, float* __restrict__ nxt, const int2* __restrict__ nb, int M, int K) {
    const long long total = (long long)M * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int id = (int)(i / K), k = (int)(i - (long long)id * K);
        const int2 n = __ldg(nb + id);
        const float a = cur[(size_t)(n.x + 1) * K + k], c = cur[(size_t)(n.y + 1) * K + k];
        nxt[(size_t)(id + 1) * K + k] = __fadd_rn(cur[(size_t)(id + 1) * K + k], __fmul_rn(0.5f, __fadd_rn(a, c)));
    }
}

// ---- 8. slice: one thread per pixel, all class planes (:536-546) ---------------------------------------------------------------
__device__ __forceinline__ void slice_phase(const float* __restrict__ values, const int* __restrict__ vertex, const float* __restrict__ weight,
                                            float* __restrict__ outs, int N, int K, int P, float alpha) {
    const long long total = (long long)N * P;
    for (long long gp = (long long)blockIdx.x * blockDim.x + threadIdx.x; gp < total; gp += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(gp / P), p = (int)(gp - (long long)b * P);
        const float* rows[PV];
        float w[PV];
#pragma unroll
        for (int r = 0; r < PV; ++r) {
            rows[r] = values + (size_t)(vertex[gp * PV + r] + 1) * K;
            w[r] = __fmul_rn(weight[gp * PV + r], alpha);
        }
        float* o = outs + (size_t)b * K * P + p;
        for (int k = 0; k < K; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < PV; ++r) acc = __fadd_rn(acc, __fmul_rn(w[r], rows[r][k]));
            o[(size_t)k * P] = acc;
        }
    }
}

__global__ void pl_splat_kernel(const float* __restrict__ insT, const unsigned* __restrict__ spix, const float* __restrict__ sweight,
                                const int* __restrict__ seg_start, const int* __restrict__ seg_end, const int* __restrict__ counters,
                                float* __restrict__ va, float* __restrict__ vb, int K) {
    splat_phase(insT, spix, sweight, seg_start, seg_end, counters[0], va, vb, K);
}
__global__ void pl_blur_kernel(const float* __restrict__ cur, float* __restrict__ nxt, const int2* __restrict__ nb,
                               const int* __restrict__ counters, int K) {
    blur_phase(cur, nxt, nb, counters[0], K);
}
__global__ void pl_slice_kernel(const float* __restrict__ values, const int* __restrict__ vertex, const float* __restrict__ weight,
                                float* __restrict__ outs, int N, int K, int P, float alpha) {
    slice_phase(values, vertex, weight, outs, N, K, P, alpha);
}

// MEASURED AND REJECTED (gpurun 2026-10-17): splat -> 6 blur passes -> slice as ONE cooperative launch with grid-wide barriers instead
// of 8 launches: 142 us vs 84 us for the separate kernels at N=2, 160x160 and 494 vs 262 us at N=8, 224x224 (64 registers x 256 threads
// hold the grid to 4 blocks/SM for every phase, and a grid barrier costs more than the kernel boundary it replaces once the call is
// replayed from a CUDA graph).  Four elements per thread and trip in the blur with L2-only loads was slower too (the rows of
// neighbouring lattice points are re-read by neighbouring threads: they want L1).

// ---- 9. element-wise part of DenseEnergyLossFunction.forward (utils/losses.py:54-64,71-74), one pass ---------------------------
// gate = 1 where unlabeled, else max(ROI - max_k seg, 0); AS *= gate; loss -= sum(seg_roi * AS) / N (double accumulation)
__global__ void dense_energy_gate_kernel(const float* __restrict__ seg, const float* __restrict__ rois, const uint8_t* __restrict__ unlabeled,
                                         const float* __restrict__ seg_roi, float* __restrict__ AS, double* __restrict__ loss_acc, int N,
                                         int K, int P) {
    double part = 0.0;
    const long long total = (long long)N * P;
    for (long long gp = (long long)blockIdx.x * blockDim.x + threadIdx.x; gp < total; gp += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(gp / P), p = (int)(gp - (long long)b * P);
        const size_t o = (size_t)b * K * P + p;
        float mx = seg[o];
        for (int k = 1; k < K; ++k) mx = fmaxf(mx, seg[o + (size_t)k * P]);
        float g = __fsub_rn(rois[gp], mx);
        if (unlabeled[gp]) g = 1.0f;
        if (g < 0.f) g = 0.f;
        for (int k = 0; k < K; ++k) {
            const float a = __fmul_rn(AS[o + (size_t)k * P], g);
            AS[o + (size_t)k * P] = a;
            part += (double)seg_roi[o + (size_t)k * P] * (double)a;
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
    __shared__ double wpart[8];
    if ((threadIdx.x & 31) == 0) wpart[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += wpart[w];
        atomicAdd(loss_acc, -t / (double)N);
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------------
struct Plan {
    long long P, P4, E, Eins, Mmax;
    unsigned cap;
    int nblk, passes, nchunks;
    size_t off_table, off_idslot, off_keyid, off_counters, off_vertex, off_weight, off_k[2], off_v[2], off_hist, off_sums, off_seg0, off_seg1, off_nb, off_insT,
        off_va, off_vb, total;
};

static inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static bool make_plan(int N, int K, int H, int W, Plan& pl) {
    if (N <= 0 || K <= 0 || K > 65535 * TK || H <= 0 || W <= 0 || N > 65535) return false;   // grid.y / grid.z of the staging kernel
    pl.P = (long long)H * W;
    pl.P4 = (pl.P + 3) & ~3ll;
    pl.E = (long long)N * pl.P * PV;
    pl.Eins = (long long)N * pl.P4 * PV;
    if (pl.Eins >= (1ll << 30)) return false;               // 32-bit entry indices, table capacity below 2^31
    pl.Mmax = pl.Eins;
    unsigned cap = 1024;
    while ((long long)cap < 2 * pl.Eins) cap <<= 1;
    pl.cap = cap;
    pl.nblk = (int)((pl.E + SORT_TILE - 1) / SORT_TILE);
    int bits = 1;
    while ((1ll << bits) < pl.Mmax) ++bits;
    pl.passes = (bits + 7) / 8;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes); return at; };
    pl.off_table = take((size_t)cap * sizeof(LKey));
    pl.off_idslot = take((size_t)cap * 4);
    pl.off_keyid = take((size_t)pl.Mmax * sizeof(LKey));
    pl.off_vertex = take((size_t)pl.E * 4);
    pl.off_weight = take((size_t)pl.E * 4);
    for (int i = 0; i < 2; ++i) { pl.off_k[i] = take((size_t)pl.E * 4); pl.off_v[i] = take((size_t)pl.E * 4); }
    pl.off_hist = take((size_t)256 * pl.nblk * 4);
    pl.nchunks = (256 * pl.nblk + SCAN_CHUNK - 1) / SCAN_CHUNK;
    pl.off_sums = take((size_t)pl.nchunks * 4);
    pl.off_insT = take((size_t)N * pl.P * K * 4);
    pl.off_counters = take(256);                            // [0] lattice points, [1 + pass] scan tickets; zeroed together with the list bounds
    pl.off_seg0 = take((size_t)(pl.Mmax + 1) * 4);
    pl.off_seg1 = take((size_t)(pl.Mmax + 1) * 4);
    pl.off_nb = take((size_t)PV * pl.Mmax * sizeof(int2));
    pl.off_va = take((size_t)(pl.Mmax + 2) * K * 4);
    pl.off_vb = take((size_t)(pl.Mmax + 2) * K * 4);
    pl.total = o;
    return true;
}

static EmbedConsts make_consts() {
    // the constants of permutohedral.cpp:124-126,156-159,535, evaluated in the reference's types
    EmbedConsts ec;
    const float inv_std_dev = (float)(sqrt(2.0 / 3.0) * (double)PV);
    for (int i = 0; i < PD; ++i) ec.scale[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * (double)inv_std_dev);
    ec.inv_v = 1.0f / (float)PV;
    ec.v = (float)PV;
    ec.alpha = 1.0f / (1.0f + powf(2.0f, -(float)PD));
    return ec;
}

static inline int grid_for(long long work, int per_block, int max_per_sm) {
    long long g = (work + per_block - 1) / per_block;
    const long long cap = (long long)num_sms() * max_per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace rss

using namespace rss;

extern "C" size_t rss_bilateral_workspace_bytes(int N, int K, int H, int W) {
    Plan pl;
    return make_plan(N, K, H, W, pl) ? pl.total : 0;
}

extern "C" int rss_bilateralfilter_batch(const float* images, const float* ins, float* outs, int N, int K, int H, int W, float sigmargb,
                                         float sigmaxy, void* workspace, size_t workspace_bytes, int* lattice_points, cudaStream_t st) {
    Plan pl;
    if (!make_plan(N, K, H, W, pl)) return RSS_ERR_SHAPE;
    if (!images || !ins || !outs || !(sigmargb > 0.f) || !(sigmaxy > 0.f)) return RSS_ERR_SHAPE;
    if (!workspace || workspace_bytes < pl.total || ((uintptr_t)workspace & 15)) return RSS_ERR_WORKSPACE;
    char* ws = (char*)workspace;
    LKey* table = (LKey*)(ws + pl.off_table);
    int* id_of_slot = (int*)(ws + pl.off_idslot);
    LKey* key_of_id = (LKey*)(ws + pl.off_keyid);
    int* counters = (int*)(ws + pl.off_counters);
    int* vertex = (int*)(ws + pl.off_vertex);
    float* weight = (float*)(ws + pl.off_weight);
    unsigned* sk[2] = {(unsigned*)(ws + pl.off_k[0]), (unsigned*)(ws + pl.off_k[1])};
    unsigned* sv[2] = {(unsigned*)(ws + pl.off_v[0]), (unsigned*)(ws + pl.off_v[1])};
    unsigned* hist = (unsigned*)(ws + pl.off_hist);
    unsigned* sums = (unsigned*)(ws + pl.off_sums);
    float* insT = (float*)(ws + pl.off_insT);
    int* seg_start = (int*)(ws + pl.off_seg0);
    int* seg_end = (int*)(ws + pl.off_seg1);
    int2* nb = (int2*)(ws + pl.off_nb);
    float* va = (float*)(ws + pl.off_va);
    float* vb = (float*)(ws + pl.off_vb);
    const EmbedConsts ec = make_consts();
    const unsigned E = (unsigned)pl.E;

    cudaMemsetAsync(table, 0xFF, (size_t)pl.cap * sizeof(LKey), st);
    // counters, scan tickets and both list bounds in one call: the arrays are adjacent up to alignment padding
    cudaMemsetAsync(counters, 0, (pl.off_seg1 - pl.off_counters) + (size_t)(pl.Mmax + 1) * 4, st);

    pl_embed_kernel<<<grid_for((long long)N * pl.P4, 256, 8), 256, 0, st>>>(images, N, (int)pl.P, (int)pl.P4, W, sigmargb, sigmaxy, ec, table,
                                                                           pl.cap - 1, id_of_slot, key_of_id, counters, vertex, weight);
    pl_entries_kernel<<<pl.nblk, 256, 0, st>>>(vertex, id_of_slot, sk[0], sv[0], E, hist, pl.nblk);
    int cur = 0;
    for (int pass = 0; pass < pl.passes; ++pass) {
        if (pass > 0) pl_sort_hist_kernel<<<pl.nblk, 256, 0, st>>>(sk[cur], E, 8 * pass, hist, pl.nblk);
        pl_scan_chunks_kernel<<<pl.nchunks, 1024, 0, st>>>(hist, 256u * (unsigned)pl.nblk, sums, counters + 1 + pass);
        pl_sort_scatter_kernel<<<pl.nblk, 256, 0, st>>>(sk[cur], sv[cur], sk[cur ^ 1], sv[cur ^ 1], E, 8 * pass, hist, sums, pl.nblk);
        cur ^= 1;
    }
    // the sort's spare buffers take the sorted pixel indices / weights
    unsigned* spix = sk[cur ^ 1];
    float* sweight = (float*)sv[cur ^ 1];
    pl_segments_kernel<<<grid_for(E, 256, 8), 256, 0, st>>>(sk[cur], sv[cur], weight, E, seg_start, seg_end, spix, sweight);
    pl_pixel_major_kernel<<<dim3((unsigned)((pl.P + TP - 1) / TP), (unsigned)N, (unsigned)((K + TK - 1) / TK)), 256, 0, st>>>(ins, insT, K, (int)pl.P);
    pl_neighbors_kernel<<<num_sms() * 8, 256, 0, st>>>(table, pl.cap - 1, id_of_slot, key_of_id, counters, nb, (int)pl.Mmax);
    pl_splat_kernel<<<num_sms() * 8, 256, 0, st>>>(insT, spix, sweight, seg_start, seg_end, counters, va, vb, K);
    float *a = va, *b = vb;
    for (int dir = 0; dir < PV; ++dir) {
        pl_blur_kernel<<<num_sms() * 8, 256, 0, st>>>(a, b, nb + (size_t)dir * pl.Mmax, counters, K);
        float* t = a; a = b; b = t;
    }
    pl_slice_kernel<<<grid_for((long long)N * pl.P, 128, 16), 128, 0, st>>>(a, vertex, weight, outs, N, K, (int)pl.P, ec.alpha);
    if (lattice_points) cudaMemcpyAsync(lattice_points, counters, sizeof(int), cudaMemcpyDeviceToDevice, st);
    return check_launch();
}

// The reference's entry point itself (bilateralfilter.hpp:12; SWIG typemaps bilateralfilter.i:21-25): HOST arrays, outs written in
// place, same argument list -- plus an int status instead of void.  The one entry point of this library that owns device memory: a
// process-wide arena grown on demand (the host caller has no device allocator to lend), copies and kernels on the legacy stream,
// synchronous like the function it replaces.  Not thread-safe (neither is the caller: the GIL is held across the SWIG call).
extern "C" int rss_bilateralfilter_batch_host(const float* images, int len_images, const float* ins, int len_ins, float* outs, int len_outs,
                                              int N, int K, int H, int W, float sigmargb, float sigmaxy) {
    Plan pl;
    if (!make_plan(N, K, H, W, pl)) return RSS_ERR_SHAPE;
    const long long n_img = (long long)N * 3 * pl.P, n_val = (long long)N * K * pl.P;
    if (len_images != n_img || len_ins != n_val || len_outs != n_val) return RSS_ERR_SHAPE;
    static char* arena = nullptr;
    static size_t arena_bytes = 0;
    const size_t io = align_up((size_t)n_img * 4) + 2 * align_up((size_t)n_val * 4);
    if (arena_bytes < io + pl.total) {
        if (arena) cudaFree(arena);
        arena = nullptr;
        arena_bytes = 0;
        if (cudaMalloc(&arena, io + pl.total) != cudaSuccess) { g_last_cuda_error = (int)cudaGetLastError(); return RSS_ERR_CUDA; }
        arena_bytes = io + pl.total;
    }
    float* d_img = (float*)arena;
    float* d_in = (float*)(arena + align_up((size_t)n_img * 4));
    float* d_out = (float*)((char*)d_in + align_up((size_t)n_val * 4));
    char* ws = (char*)d_out + align_up((size_t)n_val * 4);
    cudaMemcpyAsync(d_img, images, (size_t)n_img * 4, cudaMemcpyHostToDevice, 0);
    cudaMemcpyAsync(d_in, ins, (size_t)n_val * 4, cudaMemcpyHostToDevice, 0);
    const int rc = rss_bilateralfilter_batch(d_img, d_in, d_out, N, K, H, W, sigmargb, sigmaxy, ws, pl.total, nullptr, 0);
    if (rc != RSS_OK) return rc;
    cudaMemcpyAsync(outs, d_out, (size_t)n_val * 4, cudaMemcpyDeviceToHost, 0);
    const cudaError_t e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return RSS_ERR_CUDA; }
    return RSS_OK;
}

extern "C" int rss_dense_energy_gate(const float* seg, const float* rois, const uint8_t* unlabeled, const float* seg_roi, float* AS,
                                     double* loss_acc, int N, int K, int H, int W, cudaStream_t st) {
    if (N <= 0 || K <= 0 || H <= 0 || W <= 0 || !seg || !rois || !unlabeled || !seg_roi || !AS || !loss_acc) return RSS_ERR_SHAPE;
    const long long P = (long long)H * W;
    if (P >= (1ll << 31)) return RSS_ERR_SHAPE;
    dense_energy_gate_kernel<<<grid_for((long long)N * P, 256, 8), 256, 0, st>>>(seg, rois, unlabeled, seg_roi, AS, loss_acc, N, K, (int)P);
    return check_launch();
}
