// One-shot cross-GPU exchange for SyncBatchNorm (ffn_block.py:222,231,234: the 24 nn.SyncBatchNorm layers of the FFN) over NVLink
// peer memory, replacing the NCCL all_gather / all_reduce per layer that sat on the critical chain of the step (48 collectives of
// ~1 KB at ~20 us each = the fixed +1.1 ms of the N >= 2 step, SCALE_r01 / VERDICT r1).
//
// Every rank owns a symmetric buffer (torch.distributed._symmetric_memory: same layout on every rank, all peers mapped):
//     flags  [2 parities][world]           uint32, the sequence number of the last exchange whose data from rank r is complete
//     recv   [2 parities][world][kSyncMaxN] float
// One CTA per rank: (1) writes its vector into slot [parity][rank] of EVERY rank's recv area (peer stores over NVLink),
// (2) system-scope fence, then publishes the sequence number in every rank's flags[parity][rank], (3) spins until all `world` flags
// of its own buffer show this sequence number, (4) sums the `world` slots in rank order (every rank gets bit-identical totals).
// A buffer holds any number of such CHANNELS (chan_off): one per BatchNorm layer, each with its own sequence counter, because layers
// on different streams may reach the GPU in a different order on different ranks -- only the exchanges of ONE layer (its forward
// and its backward) are ordered the same way everywhere.
// Sequence numbers increase by one per exchange (a device-side counter, so CUDA-graph replays keep counting); two parities make the
// slot of exchange k+1 distinct from the one a slower peer may still be reading for exchange k, and a rank cannot reach exchange
// k+2 before every peer has entered k+1 (it needs their k+1 flags), so two are enough.  All ranks issue the same sequence of
// exchanges (SPMD).  The spin is bounded: a protocol error traps instead of hanging the GPU.
//
// mode 0 (backward): out[0:n] = global sums, local_out[0:n] = this rank's own vector (the parameter gradients use the local sums).
// mode 1 (forward):  vector = (sum (x-K), sum (x-K)^2 per channel, row count); the totals are turned into mean / invstd / scale /
//                    shift and the running statistics exactly like bn_finalize_channel.  Both modes clear the input accumulators.
#include "common.cuh"

namespace rss {

constexpr int kSyncMaxN = 2 * 512 + 8;          // floats per slot: 2*C + 1 for C <= 512
constexpr int kSyncMaxWorld = 16;
constexpr unsigned int kSyncSpinLimit = 1u << 27;

struct SyncBnFin {
    int C;
    const float* gamma; const float* beta;
    float* running_mean; float* running_var;
    float momentum, eps;
    float* mean_out; float* invstd_out; float* scale_out; float* shift_out;
    const float* pre_bias;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256)
sync_exchange_kernel(const int64_t* __restrict__ bases /*[world] base address of every rank's symmetric buffer*/, int64_t chan_off /*bytes*/,
                     int rank, int world,
                     unsigned int* __restrict__ counter, float* __restrict__ vec /*[n] local accumulators, cleared*/, int n,
                     float extra /*appended as element n when >= 0 (the local row count)*/, float* __restrict__ out,
                     float* __restrict__ local_out, int mode, SyncBnFin fin) {
    __shared__ unsigned int seq_s;
    __shared__ float tot[kSyncMaxN];
    if (threadIdx.x == 0) seq_s = *counter + 1u;
    __syncthreads();
    const unsigned int seq = seq_s;
    const int par = (int)(seq & 1u);
    const int nn = extra >= 0.f ? n + 1 : n;
    const size_t flag_floats = 2 * kSyncMaxWorld;                                   // flags region, in 4-byte units
    // (1) my vector -> slot [par][rank] of every rank
    for (int p = 0; p < world; ++p) {
        float* recv = reinterpret_cast<float*>(bases[p] + chan_off) + flag_floats + ((size_t)par * world + rank) * kSyncMaxN;
        for (int i = threadIdx.x; i < nn; i += blockDim.x) recv[i] = i < n ? vec[i] : extra;
    }
    __threadfence_system();
    __syncthreads();
    // (2) publish, (3) wait for everybody
    if (threadIdx.x < world) {
        unsigned int* pf = reinterpret_cast<unsigned int*>(bases[threadIdx.x] + chan_off) + (size_t)par * kSyncMaxWorld + rank;
        st_release_sys(pf, seq);
        const unsigned int* mf = reinterpret_cast<const unsigned int*>(bases[rank] + chan_off) + (size_t)par * kSyncMaxWorld + threadIdx.x;
        unsigned int spins = 0;
        while (ld_acquire_sys(mf) != seq) {
            if (++spins > kSyncSpinLimit) __trap();
        }
    }
    __syncthreads();
    // (4) totals, summed in rank order
    const float* mine = reinterpret_cast<const float*>(bases[rank] + chan_off) + flag_floats + (size_t)par * world * kSyncMaxN;
    for (int i = threadIdx.x; i < nn; i += blockDim.x) {
        float t = 0.f;
        for (int r = 0; r < world; ++r) t += __ldcv(mine + (size_t)r * kSyncMaxN + i);
        tot[i] = t;
    }
    __syncthreads();
    if (mode == 0) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            out[i] = tot[i];
            if (local_out) local_out[i] = vec[i];
            vec[i] = 0.f;
        }
    } else {
        const int C = fin.C;
        const float ng = tot[n];                                                  // global row count
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const float S = tot[c], Q = tot[C + c];
            const float md = S / ng, m2 = fmaxf(Q - S * md, 0.f);
            const float k = fin.running_mean ? fin.running_mean[c] - (fin.pre_bias ? fin.pre_bias[c] : 0.f) : 0.f;
            const float mean = k + md, invstd = rsqrtf(m2 / ng + fin.eps);
            fin.mean_out[c] = mean;
            fin.invstd_out[c] = invstd;
            const float sc = fin.gamma[c] * invstd;
            fin.scale_out[c] = sc;
            fin.shift_out[c] = fin.beta[c] - mean * sc;
            if (fin.running_mean) {
                fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * (mean + (fin.pre_bias ? fin.pre_bias[c] : 0.f));
                fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * (m2 / fmaxf(ng - 1.f, 1.f));
            }
            vec[c] = 0.f;
            vec[C + c] = 0.f;
        }
    }
    if (threadIdx.x == 0) *counter = seq;
}

}  // namespace rss

using namespace rss;

extern "C" size_t rss_sync_exchange_bytes(int world) {
    if (world < 1 || world > kSyncMaxWorld) return 0;
    return (2 * (size_t)kSyncMaxWorld + 2 * (size_t)world * kSyncMaxN) * sizeof(float);
}

// backward exchange: out[0:n] = sum over ranks of vec, local_out[0:n] = vec (may be NULL); vec is cleared
extern "C" int rss_sync_allreduce_small(const int64_t* bases, int64_t chan_off, int rank, int world, unsigned int* counter, float* vec, int n,
                                        float* out, float* local_out, cudaStream_t st) {
    if (chan_off < 0 || (chan_off & 15) || !bases || !counter || !vec || !out || n <= 0 || n > kSyncMaxN - 1 || world < 1 || world > kSyncMaxWorld || rank < 0 || rank >= world)
        return RSS_ERR_SHAPE;
    SyncBnFin fin{};
    sync_exchange_kernel<<<1, 256, 0, st>>>(bases, chan_off, rank, world, counter, vec, n, -1.f, out, local_out, 0, fin);
    return check_launch();
}

// forward exchange + finalize: accum [2C] = this rank's sum (x-K), sum (x-K)^2 (rss_bn_stats_raw), local_rows its row count
extern "C" int rss_sync_bn_finalize(const int64_t* bases, int64_t chan_off, int rank, int world, unsigned int* counter, float* accum, int C,
                                    int64_t local_rows, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                    float momentum, float eps, float* mean_out, float* invstd_out, float* scale, float* shift,
                                    const float* pre_bias, cudaStream_t st) {
    if (chan_off < 0 || (chan_off & 15) || !bases || !counter || !accum || C <= 0 || 2 * C > kSyncMaxN - 1 || world < 1 || world > kSyncMaxWorld || rank < 0 || rank >= world ||
        local_rows <= 0)
        return RSS_ERR_SHAPE;
    SyncBnFin fin;
    fin.C = C; fin.gamma = gamma; fin.beta = beta; fin.running_mean = running_mean; fin.running_var = running_var; fin.momentum = momentum;
    fin.eps = eps; fin.mean_out = mean_out; fin.invstd_out = invstd_out; fin.scale_out = scale; fin.shift_out = shift; fin.pre_bias = pre_bias;
    sync_exchange_kernel<<<1, 256, 0, st>>>(bases, chan_off, rank, world, counter, accum, 2 * C, (float)local_rows, nullptr, nullptr, 1, fin);
    return check_launch();
}
