// Evaluation-side device kernels: the pieces of RSSFormer-TIP2023/train.py:17,42-55 / eval.py:48-80 /
// module/tta.py:12-24,118-137 that the reference runs on the host or through generic torch ops.
//   * rss_bilinear_resize: F.interpolate(mode='bilinear', align_corners=True) of NCHW fp32 planes with an axpby epilogue
//     (dst = alpha*resize(src) + beta*dst): the Scale transform of the test-time augmentation and its inverse, the latter
//     accumulating the running mean of the per-scale probabilities in place (tta.py:18-22: sum(outs) / len(outs));
//   * rss_confusion_matrix: PixelMetric.forward (train.py:48-49): counts[truth][prediction] over the valid pixels, from the uint8
//     arg-max map rss_head_probs already produces -- the (B,7,H,W) probabilities never travel to the host.
#include "common.cuh"

namespace rss {

__global__ void bilinear_resize_kernel(const float* __restrict__ src, float* __restrict__ dst, int planes, int h, int w, int H, int W,
                                       float sy, float sx, float alpha, float beta) {
    const int64_t total = (int64_t)planes * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int X = (int)(i % W), Y = (int)((i / W) % H);
        const int64_t pl = i / ((int64_t)W * H);
        // align_corners=True: source coordinate = destination index * (in - 1) / (out - 1)
        const float fy = sy * (float)Y, fx = sx * (float)X;
        int y0 = (int)fy, x0 = (int)fx;
        if (y0 > h - 1) y0 = h - 1;
        if (x0 > w - 1) x0 = w - 1;
        const int y1 = y0 < h - 1 ? y0 + 1 : y0, x1 = x0 < w - 1 ? x0 + 1 : x0;
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        const float* p = src + pl * (int64_t)h * w;
        const float v = (1.f - ly) * ((1.f - lx) * __ldg(p + (int64_t)y0 * w + x0) + lx * __ldg(p + (int64_t)y0 * w + x1)) +
                        ly * ((1.f - lx) * __ldg(p + (int64_t)y1 * w + x0) + lx * __ldg(p + (int64_t)y1 * w + x1));
        dst[i] = beta == 0.f ? alpha * v : alpha * v + beta * dst[i];
    }
}

constexpr int kCmMax = 16;
__global__ void confusion_matrix_kernel(const uint8_t* __restrict__ pred, const int64_t* __restrict__ truth, unsigned long long* __restrict__ cm,
                                        int64_t n, int K, int ignore_index) {
    __shared__ unsigned int hist[kCmMax * kCmMax];
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = truth[i];
        const int p = pred[i];
        if (t != ignore_index && t >= 0 && t < K && p < K) atomicAdd(&hist[(int)t * K + p], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += blockDim.x)
        if (hist[i]) atomicAdd(cm + i, (unsigned long long)hist[i]);
}

}  // namespace rss

using namespace rss;

extern "C" int rss_bilinear_resize(const float* src, float* dst, int planes, int h, int w, int H, int W, float alpha, float beta,
                                   cudaStream_t st) {
    if (planes <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return RSS_ERR_SHAPE;
    const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const int64_t total = (int64_t)planes * H * W;
    int grid = (int)((total + 255) / 256);
    if (grid > num_sms() * 16) grid = num_sms() * 16;
    bilinear_resize_kernel<<<grid, 256, 0, st>>>(src, dst, planes, h, w, H, W, sy, sx, alpha, beta);
    return check_launch();
}

extern "C" int rss_confusion_matrix(const uint8_t* pred, const int64_t* truth, unsigned long long* cm_acc, int64_t n, int num_classes,
                                    int ignore_index, cudaStream_t st) {
    if (n <= 0 || num_classes <= 0 || num_classes > kCmMax) return RSS_ERR_SHAPE;
    int grid = (int)((n + 256 * 16 - 1) / (256 * 16));
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    if (grid < 1) grid = 1;
    confusion_matrix_kernel<<<grid, 256, 0, st>>>(pred, truth, cm_acc, n, num_classes, ignore_index);
    return check_launch();
}
