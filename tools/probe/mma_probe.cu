// Micro-probe: cycles per tcgen05.mma (kind::f16, bf16, cta_group::1, M=128, SS operands, K-major SWIZZLE_128B) as a function of
// N, of the number of independent accumulators, and of the A start-address alignment (1024-byte aligned vs shifted by whole
// 128-byte rows, as conv_cf.cu / conv_wgrad_tc.cu do for their tap shifts).  One CTA per SM, 256 back-to-back MMAs per measurement.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/mma_probe tools/probe/mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../representationlearning_b200/csrc/tc05.cuh"
using namespace rss;
namespace rss { int g_last_cuda_error = 0; }

__device__ __forceinline__ uint32_t idesc_of(int n, int mn_major) {
    uint32_t d = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (mn_major) d |= (1u << 15) | (1u << 16);
    return d;
}

// mode bits: 0 = aligned K-major; 1 = A start shifted by (i % 3) rows; 2 = MN-major operands
__global__ void __launch_bounds__(128, 1) probe(int N, int nacc, int mode, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&slot), 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 48 * 1024;
        const uint32_t idesc = idesc_of(N, mode & 2);
        uint64_t hi;
        if (mode & 2) { hi = 0; hi |= (uint64_t)((8 * 1024) >> 4) << 16; hi |= (uint64_t)(1024 >> 4) << 32; hi |= (uint64_t)1 << 46; hi |= (uint64_t)2 << 61; }
        else hi = make_sw128_desc_bo(0, 0);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t sh = (mode & 1) ? (uint32_t)(i % 3) * 8 : 0u;          // +0/1/2 rows of 128 B
            const uint32_t kk = (uint32_t)(i & 3) * 2;                             // the 4 K=16 steps of a 128-byte row
            const uint32_t acc = (uint32_t)(i % nacc) * N;
            umma_bf16(tmem + acc, hi | (uint64_t)((a0 >> 4) + sh + ((mode & 2) ? 0 : kk)), hi | (uint64_t)((b0 >> 4) + ((mode & 2) ? 0 : kk)),
                      idesc, i >= nacc);
        }
        umma_commit(smem_u32(&bar));
        long long t1 = clock64();
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        out[blockIdx.x * 2] = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * 2 * sizeof(long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 256;
    printf("# cycles per tcgen05.mma M=128 K=16 bf16 SS (issue = time to issue all, total = until commit arrives), %d MMAs, grid 148\n", iters);
    printf("%6s %5s %5s %10s %10s\n", "N", "nacc", "mode", "issue/mma", "total/mma");
    const int Ns[] = {32, 64, 128, 256};
    for (int mode = 0; mode < 4; ++mode)
        for (int ni = 0; ni < 4; ++ni)
            for (int nacc = 1; nacc <= 4; nacc *= 2) {
                const int N = Ns[ni];
                if (nacc * N > 512) continue;
                for (int rep = 0; rep < 2; ++rep) probe<<<148, 128, 100 * 1024>>>(N, nacc, mode, iters, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("N=%d nacc=%d mode=%d: %s\n", N, nacc, mode, cudaGetErrorString(e)); return 1; }
                long long h[296];
                cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                double a = 0, b = 0;
                for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
                printf("%6d %5d %5d %10.1f %10.1f\n", N, nacc, mode, a / 148 / iters, b / 148 / iters);
            }
    return 0;
}
