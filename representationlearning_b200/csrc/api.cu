#include "common.cuh"

namespace rss { int g_last_cuda_error = 0; }

extern "C" int rss_version(void) { return 100; }
extern "C" int rss_last_cuda_error(void) { return rss::g_last_cuda_error; }

extern "C" int rss_check_device(void) {
    int dev = 0, major = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) { rss::g_last_cuda_error = (int)e; return RSS_ERR_CUDA; }
    return major == 10 ? RSS_OK : RSS_ERR_ARCH;
}
