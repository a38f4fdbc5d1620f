/* rss_b200.h — C ABI of the B200-native (sm_100a) RSSFormer training hot path.
 *
 * The reference (Rongtao-Xu/RepresentationLearning, RSSFormer-TIP2023) has NO FFI on this path: the
 * seam it offers is the Python nn.Module surface.  This header therefore defines the
 * boundary a maintainer binds from those modules (ctypes stub shown in INTEGRATION.md); every entry
 * point names the reference code it replaces.  The only FFI precedent in the reference tree is the
 * SWIG signature `void bilateralfilter_batch(float* images,int, float* ins,int, float* outs,int, int N,
 * int K,int H,int W, float,float)` (SCD-AAAI2023/wrapper/bilateralfilter/bilateralfilter.hpp:12,
 * bilateralfilter.i:21-25): caller-allocated flat arrays, in-place outputs.  The same conventions hold
 * here, plus an int status:
 *
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's caching allocator); the library
 *     allocates nothing and launches only on the `stream` it is given (one exception, stated at its declaration:
 *     rss_bilateralfilter_batch_host takes HOST arrays and owns a device arena); the only state kept between calls is per-process
 *     launch configuration read once (opt-in shared-memory attributes, SM count, RSS_* environment switches): one GPU per process;
 *   - no host synchronisation, no host reads of device data;
 *   - activations are NHWC (== token-major (B, H*W, C)), element type selected by `dtype`
 *     (RSS_F32 or RSS_BF16); parameters, statistics and all accumulation are fp32;
 *   - return value: RSS_OK (0) or a negative rss_status; the CUDA error code of a failed launch is kept
 *     in rss_last_cuda_error().  Nothing throws or exits;
 *   - `*_acc` gradient outputs are ACCUMULATED into (caller zeroes them once per step).
 *
 * Paths below are relative to RSSFormer-TIP2023/module/baseline/ in the reference tree.
 */
#ifndef RSS_B200_H_
#define RSS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

typedef enum { RSS_OK = 0, RSS_ERR_SHAPE = -1, RSS_ERR_DTYPE = -2, RSS_ERR_CUDA = -3, RSS_ERR_WORKSPACE = -4,
               RSS_ERR_ARCH = -5 } rss_status;
typedef enum { RSS_F32 = 0, RSS_BF16 = 1 } rss_dtype;
typedef enum { RSS_ACT_NONE = 0, RSS_ACT_RELU = 1, RSS_ACT_GELU = 2 } rss_act;

int rss_version(void);
int rss_last_cuda_error(void);
/* 0 when the current device is sm_100 (B200); RSS_ERR_ARCH otherwise. There is no fallback path. */
int rss_check_device(void);

/* ---- LayerNorm over channels of (rows, C) tokens: base_hrnet/modules/MTFM.py:64,78-79,107,109 ---- */
int rss_layernorm_fwd(const void* x, void* y /*NULL: statistics only*/, float* mean /*[rows] or NULL*/, float* rstd /*[rows]*/,
                      const float* gamma, const float* beta, float eps, int64_t rows, int C, int dtype, cudaStream_t stream);
/* dx = LN'(dy) (+ dx_add if not NULL); dgamma_acc/dbeta_acc accumulate */
int rss_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                      const void* dx_add, void* dx, float* dgamma_acc, float* dbeta_acc,
                      int64_t rows, int C, int dtype, cudaStream_t stream);

/* ---- gate + window cross-attention region: MTFM.py:101-107 (norm1 on both inputs, residual),
 *      base_hrnet/modules/multihead_isa_pool_attention.py:148-188 (InterlacedPoolAttention2.forward),
 *      multihead_isa_attention.py:373-426 (PadBlock, LocalPermuteModule), DAL.py:785-1030 (Mhca) ---- */
typedef struct {
    const float *ln_w, *ln_b;                  /* transformer.norm1 (C)                        */
    const float *sa1_w, *sa2_w;                /* attn.atrous_block{1,2}.conv1.weight (1,2,7,7) */
    const float *lvl_w, *lvl_b;                /* attn.weight_levels (2,2,1,1), (2)            */
    const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b;   /* attn.attn.{q,k,v,out}_proj (C,C),(C) */
    float ln_eps;
    int C, num_heads, window;                  /* must be 32, 2, 7 (the only geometry the reference builds: _hrnet_rssformer.py:308) */
} rss_attn_params;
typedef struct {
    float *ln_w, *ln_b, *sa1_w, *sa2_w, *lvl_w, *lvl_b, *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b;
} rss_attn_grads;

/* out = x + Attn(LN1(x), LN1(y)).  Saved for backward (caller-owned): ln_stats [4][B*HW] f32 (mean/rstd of x, of y);
 * pooled [B][4][HW] f32; amax [B][2][HW] u8; smap, gmap [B][2][HW] f32.  The normalised tokens are never materialised:
 * every consumer recomputes (x-mean)*rstd*gamma+beta in fp32 from x and ln_stats. */
#define RSS_ATTN_SIMT 2          /* flags bit 1: force the fp32 SIMT kernels even for bf16 activations (A/B testing) */
#define RSS_ATTN_NO_GATE 4       /* flags bit 2: no saliency gate (Mhca.forward alone, DAL.py:726-735): gmap is INPUT, filled by the caller */
#define RSS_ATTN_NO_RESIDUAL 1   /* flags bit 0: out = Attn(...) without "+ x" (InterlacedPoolAttention2.forward alone) */
/* p->ln_w == NULL skips norm1 (x, y are then the already-normalised tokens; ln_stats unused). */
/* SpatialAttention.forward alone (multihead_isa_pool_attention.py:101-115): out (B,1,H,W) f32 = sigmoid(conv7x7([mean_c, max_c]))
 * of an NCHW-contiguous (B,32,H,W) tensor.  Workspaces: pooled [B][4][HW] f32, amax [B][2][HW] u8, maps [2][B][2][HW] f32. */
int rss_spatial_attention_fwd(const void* x_nchw, const float* conv_w, float* out, float* ws_pooled, uint8_t* ws_amax, float* ws_maps,
                              int B, int H, int W, int dtype, cudaStream_t stream);
int rss_attn_fwd(const void* x, const void* y, const rss_attn_params* p, int B, int H, int W, int dtype, int flags,
                 float* ln_stats, float* pooled, uint8_t* amax, float* smap, float* gmap,
                 void* out, cudaStream_t stream);
size_t rss_attn_bwd_workspace_bytes(int B, int H, int W, int dtype);
int rss_attn_bwd(const void* dout, const void* x, const void* y, const rss_attn_params* p, int B, int H, int W, int dtype, int flags,
                 const float* ln_stats, const float* pooled, const uint8_t* amax,
                 const float* smap, const float* gmap, void* workspace, size_t workspace_bytes,
                 void* dx, void* dy, const rss_attn_grads* grads_acc, cudaStream_t stream);

/* ---- BatchNorm / SyncBatchNorm (training) fused with the following activation and residual add:
 *      _hrnet_rssformer.py:216-287 (BasicBlock/Bottleneck), ffn_block.py:222,231,234,246-263 (MlpDWBN) ---- */
int rss_bn_stats_nparts(int64_t rows, int C);          /* number of partials rss_bn_stats writes */
int rss_bn_stats(const void* x, float* partials /*[nparts][C][2] (mean,M2)*/, float* counts /*[nparts]*/,
                 int64_t rows, int C, int dtype, cudaStream_t stream);
/* single-GPU path: statistics + finalize in ONE launch.  accum_scratch: persistent float[2*C] (>= 4096 floats recommended),
 * ticket: persistent uint; both zero before the first call and left zero by every call; calls sharing them must be
 * stream-ordered.  The running mean (if given) is used as the shift of the one-pass variance. */
int rss_bn_stats_fused(const void* x, float* accum_scratch, unsigned int* ticket, int64_t rows, int C, int dtype,
                       const float* gamma, const float* beta, float* running_mean /*may be NULL*/, float* running_var,
                       float momentum, float eps, float* mean_out, float* invstd_out, float* scale, float* shift,
                       const float* pre_bias /*may be NULL: bias of the producing conv that was NOT added to x (it cancels in the
                                               normalisation; only the running mean sees it)*/,
                       cudaStream_t stream);
int rss_bn_combine(const float* partials, const float* counts, int nparts, int C, float* stat /*[C][2]*/, float* total /*[1]*/,
                   cudaStream_t stream);
int rss_bn_finalize(const float* stat, const float* total, const float* gamma, const float* beta,
                    float* running_mean /*may be NULL*/, float* running_var, float momentum, float eps, int C,
                    float* mean_out, float* invstd_out, float* scale, float* shift, const float* pre_bias /*may be NULL*/,
                    cudaStream_t stream);
int rss_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float eps, int C, float* mean_out, float* invstd_out, float* scale, float* shift, cudaStream_t stream);
int rss_bn_act_fwd(const void* x, const void* residual /*may be NULL*/, void* y, const float* scale, const float* shift,
                   int64_t rows, int C, int act, int dtype, cudaStream_t stream);
/* sums[0..C) = sum dz, sums[C..2C) = sum dz*xhat with dz = dy*act'(.) (== dbeta, dgamma).  y (the saved output) is needed only
 * for RELU layers that had a residual; with y == NULL the ReLU mask is recomputed from x*scale+shift. */
int rss_bn_bwd_reduce(const void* x, const void* y, const void* dy, const float* scale, const float* shift,
                      const float* mean, const float* invstd, float* sums, int64_t rows, int C, int act, int dtype,
                      cudaStream_t stream);
/* same, without the memset node in front of the kernel: accum_scratch (float[2C]) + ticket (one counter) are a persistent
 * per-layer scratch, zero on entry and left zero (the last block to arrive publishes the totals into `sums`); both may be
 * NULL (then `sums` is cleared by a memset node first); accum_scratch without ticket = "raw" protocol (totals stay in the scratch,
 * `sums` is not written, see rss_bn_bwd_apply_raw). */
int rss_bn_bwd_reduce_ws(const void* x, const void* y, const void* dy, const float* scale, const float* shift,
                         const float* mean, const float* invstd, float* sums, float* accum_scratch, unsigned int* ticket,
                         void* dz_out /*may be NULL: [rows][C] activation dtype, receives dz for rss_bn_bwd_apply_dz*/,
                         int64_t rows, int C, int act, int dtype, cudaStream_t stream);
/* "Raw" protocol of the single-rank training path: rss_bn_stats_raw only adds per-block sums of (x-K), (x-K)^2 (K = running
 * mean - pre_bias) into the layer's persistent zeroed scratch; rss_bn_act_fwd_raw derives scale/shift from those totals in its
 * prologue, and its last block writes mean/invstd/scale/shift, updates the running statistics and clears the scratch + ticket.
 * Same results as rss_bn_stats_fused + rss_bn_act_fwd with three dependent global round trips less per layer.
 * Backward: rss_bn_bwd_reduce_ws(accum_scratch, ticket = NULL) + rss_bn_bwd_apply_raw (reads the totals from the scratch; its last
 * block publishes sums_out (optional), accumulates dgamma/dbeta and clears the scratch). */
int rss_bn_stats_raw(const void* x, float* accum_scratch, int64_t rows, int C, int dtype,
                     const float* running_mean /*may be NULL*/, const float* pre_bias /*may be NULL*/, cudaStream_t stream);
int rss_bn_act_fwd_raw(const void* x, const void* residual /*may be NULL*/, void* y, float* accum_scratch, unsigned int* ticket,
                       int64_t rows, int C, int act, int dtype, const float* gamma, const float* beta,
                       float* running_mean /*may be NULL*/, float* running_var, float momentum, float eps,
                       float* mean_out, float* invstd_out, float* scale, float* shift, const float* pre_bias /*may be NULL*/,
                       cudaStream_t stream);
int rss_bn_bwd_apply_raw(const void* x, const void* y, const void* dy, const float* scale, const float* shift,
                         const float* mean, const float* invstd, float* accum_scratch, unsigned int* ticket,
                         float inv_count, void* dx, void* dresidual /*may be NULL*/, int64_t rows, int C, int act, int dtype,
                         float* sums_out /*may be NULL*/, float* dgamma_acc /*may be NULL*/, float* dbeta_acc, cudaStream_t stream);
/* second pass from the stored dz (used for the GELU layers, whose derivative is too expensive to recompute in both passes) */
int rss_bn_bwd_apply_dz(const void* x, const void* dz, const float* scale, const float* mean, const float* invstd,
                        const float* sums, float inv_count, void* dx, int64_t rows, int C, int dtype,
                        const float* local_sums, float* dgamma_acc /*may be NULL*/, float* dbeta_acc, cudaStream_t stream);
int rss_bn_bwd_apply(const void* x, const void* y, const void* dy, const float* scale, const float* shift,
                     const float* mean, const float* invstd, const float* sums, float inv_count,
                     void* dx, void* dresidual /*may be NULL: receives dz*/, int64_t rows, int C, int act, int dtype,
                     const float* local_sums /*this rank's sums (== sums without SyncBN)*/,
                     float* dgamma_acc /*may be NULL*/, float* dbeta_acc, cudaStream_t stream);

/* ---- SyncBatchNorm exchange over NVLink peer memory (ffn_block.py:222,231,234), replacing one NCCL collective per layer and pass.
 * `bases[r]` = address of rank r's symmetric buffer (channels of rss_sync_exchange_bytes(world) bytes each, zero-initialised, every
 * peer mapped, e.g. torch.distributed._symmetric_memory); one CHANNEL (chan_off_bytes) and one zero-initialised device uint32
 * `counter` per BatchNorm layer: only the exchanges of one layer are ordered identically on all ranks.  One-CTA kernels: write the
 * local vector into every rank's slot, system fence, publish a sequence number, spin (bounded) for all ranks, sum in rank order.
 * All ranks must issue the same sequence of exchange calls.
 *   rss_sync_allreduce_small: out[0:n] = sum over ranks of vec, local_out = vec (may be NULL); vec is cleared (backward sums).
 *   rss_sync_bn_finalize:     accum[2C] = this rank's raw statistics (rss_bn_stats_raw) -> global mean / invstd / scale / shift and
 *                             running statistics (as rss_bn_finalize); accum is cleared. */
size_t rss_sync_exchange_bytes(int world);
int rss_sync_allreduce_small(const int64_t* bases, int64_t chan_off_bytes, int rank, int world, unsigned int* counter, float* vec, int n,
                             float* out, float* local_out, cudaStream_t stream);
int rss_sync_bn_finalize(const int64_t* bases, int64_t chan_off_bytes, int rank, int world, unsigned int* counter, float* accum, int C,
                         int64_t local_rows,
                         const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                         float* mean_out, float* invstd_out, float* scale, float* shift, const float* pre_bias, cudaStream_t stream);


/* ---- multi-resolution fuse sum of HighResolutionModule.forward (_hrnet_rssformer.py:418-435) and the residual+ReLU closing a
 *      transformer block (MTFM.py:109, _hrnet_rssformer.py:435): out = [relu](sum_j nearest_up_{2^k_j}(term_j)), NHWC ---- */
int rss_fuse_sum_fwd(const void* const* terms, const int* log2_up, int n_terms /*1..4*/, void* out, int B, int H, int W, int C,
                     int relu, int dtype, cudaStream_t stream);
/* gradient of ONE term: block-sum over its 2^k x 2^k footprint of dout * [out > 0 if relu] */
int rss_fuse_sum_bwd(const void* dout, const void* out /*needed if relu*/, void* dterm, int log2_up, int B, int H, int W, int C,
                     int relu, int dtype, cudaStream_t stream);

/* ---- tcgen05/TMA implicit-GEMM convolution (stride 1, "same" padding, NHWC bf16, fp32 TMEM accumulation):
 *      the FFN's dw(1x1)+dw6(3x3,d6)+dw12(3x3,d12) summed convs as ONE 17-tap GEMM (ffn_block.py:226-228,250-257),
 *      fc1/fc2 (ffn_block.py:219,232), the neck 1x1 (hrnet_aux.py:46), HRNet stride-1 3x3/1x1 convs
 *      (_hrnet_rssformer.py:209-213,253-259) and their data gradients (transposed pack, negated taps). ---- */
int rss_conv_igemm_supported(int B, int H, int W, int Cin, int Cout);
size_t rss_conv_packed_bytes(int n_srcs, const int* ksizes, int Cout, int Cin);
/* packs 1..3 parallel convs (fp32 (Cout,Cin,k,k), k in {1,3}, summed outputs) into bf16 [tap][N][K]; coinciding taps are
 * merged; bias_sum[Cout] = sum of the biases (forward pack only); taps_*_out sized 32.  transpose=1: data-gradient operand. */
int rss_conv_pack_weights(const float* const* weights, const float* const* biases, const int* ksizes, const int* dilations,
                          int n_srcs, int Cout, int Cin, int transpose, void* packed, float* bias_sum,
                          int* n_taps_out, int* taps_dy_out, int* taps_dx_out, cudaStream_t stream);
int rss_conv_igemm(const void* x, const void* w_packed, const float* bias /*may be NULL*/, void* y, int B, int H, int W,
                   int Cin, int Cout, int n_taps, const int* taps_dy, const int* taps_dx, cudaStream_t stream);
/* the same GEMM with a BatchNorm-statistics epilogue (ffn_block.py:250-258: dw + dw6 + dw12 -> norm2): besides y, adds per output
 * channel c sum (y - K_c) into stat_accum[c] and sum (y - K_c)^2 into stat_accum[Cout + c] over the bf16-rounded outputs (K =
 * stat_shift - stat_shift_sub, each [Cout] or NULL = 0: the layer's running mean, minus the bias of the producing convolution when
 * that bias is left out of y (fc1 -> norm1, fc2 -> norm3: ffn_block.py:246-263); keeps E[d^2] - E[d]^2 well conditioned) -- the "raw sums" contract of
 * rss_bn_stats_raw, consumed by rss_bn_act_fwd_raw / rss_sync_bn_finalize, so no statistics pass reads the tensor again.
 * stat_accum: the layer's persistent zeroed scratch (the consumer clears it); Cout <= 128. */
int rss_conv_igemm_stats(const void* x, const void* w_packed, const float* bias /*may be NULL*/, void* y, int B, int H, int W,
                         int Cin, int Cout, int n_taps, const int* taps_dy, const int* taps_dx, float* stat_accum,
                         const float* stat_shift /*may be NULL*/, const float* stat_shift_sub /*may be NULL*/, cudaStream_t stream);

/* ---- fused tcgen05 convolution of the HRNet family (BasicBlock/Bottleneck 3x3 and 1x1 stride-1 convs and their data gradients:
 *      _hrnet_rssformer.py:209-287):  y = E(conv(T(x)) + add)
 *        T = identity or relu(x*in_scale + in_shift): the previous layer's BatchNorm(+ReLU) applied to the staged tile;
 *        E = RSS_CF_PLAIN  identity
 *            RSS_CF_STATS  + training-mode BatchNorm statistics of y with the contract of rss_bn_stats_fused (persistent zeroed
 *                          accum[2*Cout] + ticket; outputs mean/invstd/scale/shift; running statistics updated)
 *            RSS_CF_BNRED  y is the gradient w.r.t. the output of a BatchNorm(+ReLU) with input bn_z: stores g = y * relu_mask
 *                          (mask = bn_out > 0 when bn_out is given, else bn_scale*bn_z + bn_shift > 0 when bn_relu, else 1)
 *                          and sums_out[0:Cout] = sum g, sums_out[Cout:2*Cout] = sum g * (bn_z - bn_mean) * bn_invstd
 *                          (what rss_bn_bwd_apply needs; same accum/ticket contract).
 *      Every input pixel is staged in shared memory once per tile (TMA box of whole zero-padded rows, 128B-swizzled K-major
 *      UMMA layout) and all taps read it at shifted descriptor addresses.  w_packed: bf16 [tap][Cout][Cin] from
 *      rss_conv_pack_weights (transpose=1 pack with Cin/Cout swapped gives the data gradient), or any layout with contiguous
 *      input channels described by the two strides -- e.g. the per-step channels-last / transposed shadows of
 *      rss_shadow_cl_refresh / rss_shadow_t_refresh, so no pack kernel runs per call. ---- */
#define RSS_CF_PLAIN 0
#define RSS_CF_STATS 1
#define RSS_CF_BNRED 2
typedef struct {
    int mode;                          /* RSS_CF_* */
    const void* add;                   /* bf16 (B,H,W,Cout) added before E, or NULL (residual gradient / residual) */
    float* accum;                      /* STATS, BNRED: persistent zeroed [2*Cout] */
    unsigned int* ticket;              /* STATS, BNRED: persistent zeroed */
    const float* gamma; const float* beta;             /* STATS */
    float* running_mean; float* running_var;           /* STATS, may be NULL */
    float momentum, eps;                               /* STATS */
    float* mean_out; float* invstd_out; float* scale_out; float* shift_out;   /* STATS */
    const void* bn_z; const void* bn_out;              /* BNRED (bn_out may be NULL) */
    const float* bn_mean; const float* bn_invstd; const float* bn_scale; const float* bn_shift;   /* BNRED */
    int bn_relu;                                       /* BNRED */
    float* sums_out;                                   /* BNRED: [2*Cout] */
} RssConvCfEpilogue;
int rss_conv_cf_supported(int B, int H, int W, int Cin, int Cout, int ksize, int mode);
int rss_conv_cf(const void* x, const void* w_packed, void* y, int B, int H, int W, int Cin, int Cout,
                int n_taps, const int* taps_dy, const int* taps_dx,
                int w_row_stride, int w_tap_stride /* elements; weight (tap,co,ci) at co*row + tap*tap_stride + ci; 0,0 = packed */,
                const float* in_scale /*[Cin] or NULL*/, const float* in_shift, int in_relu,
                const RssConvCfEpilogue* epilogue /*NULL = plain*/, cudaStream_t stream);

/* ---- weight gradient of the HRNet-family convolutions (BasicBlock/Bottleneck 3x3, stride-2 3x3 of the transition/fuse layers,
 *      small 1x1 convs: _hrnet_rssformer.py:209-287,361-405,512-546; FFN fc1/fc2: ffn_block.py:219,232).  Replaces the
 *      weight half of torch's conv backward.  x (B,Hi,Wi,Cin), dy (B,Ho,Wo,Cout) bf16 NHWC; dw_acc (Cout,Cin,k,k) fp32 is
 *      ACCUMULATED into with float atomics (split-K over CTAs), so it may be the optimiser's flat gradient buffer. ---- */
int rss_conv_wgrad_supported(int Cin, int Cout, int ksize, int stride, int pad, int dil);
int rss_conv_wgrad(const void* x, const void* dy, float* dw_acc, int B, int Hi, int Wi, int Cin, int Ho, int Wo, int Cout,
                   int ksize, int stride, int pad, int dil, cudaStream_t stream);

/* tcgen05 version for the stride-1 layers (3x3 pad 1, 1x1): both operands staged by TMA as zero-padded rows in the MN-major
 * 128B-swizzled UMMA layout (positions = K), taps = row shifts of the X operand, accumulators in TMEM, split-K over CTAs.
 * in_scale/in_shift/in_relu: optional relu(x*scale+shift) applied to x on load (BatchNorm+ReLU of a tensor never materialised). */
int rss_conv_wgrad_tc_supported(int B, int H, int W, int Cin, int Cout, int ksize);
int rss_conv_wgrad_tc(const void* x, const void* dy, float* dw_acc, int B, int H, int W, int Cin, int Cout, int ksize,
                      const float* in_scale /*[Cin] or NULL*/, const float* in_shift, int in_relu, cudaStream_t stream);

/* ---- neck: hrnet_aux.py:51-68 (SimpleFusion8: 3x bilinear align_corners=True + concat), NHWC ---- */
int rss_neck_gather_fwd(const void* f0, const void* f1, const void* f2, const void* f3, void* out_cat,
                        int B, const int* C /*[4]*/, const int* h /*[4]*/, const int* w /*[4]*/, int dtype, cudaStream_t stream);
int rss_neck_gather_bwd(const void* dcat, void* d0, void* d1, void* d2, void* d3,
                        int B, const int* C, const int* h, const int* w, int dtype, cudaStream_t stream);

/* ---- HRNet stem, first convolution: _hrnet_rssformer.py:467-470 Conv2d(3,64,3,stride 2,padding 1,bias=False) read straight
 *      from the (B,3,H,W) planar fp32 (or bf16) image batch the reference model receives (csrc/stem.cu): the cast, the
 *      NCHW->NHWC change, the convolution and -- optionally -- the BatchNorm raw sums of bn1 in one launch.
 *      y: (B,Ho,Wo,64) bf16 NHWC, Ho=(H-1)/2+1, Wo=(W-1)/2+1.  stat_accum (NULL or [2][64] fp32, zero or holding earlier partial
 *      sums) += sum (y-K), sum (y-K)^2 over the bf16-rounded outputs, K = stat_shift[c] (NULL: 0): the operand of
 *      rss_bn_act_fwd_raw.  No data gradient (the image needs none); dw_acc (64,3,3,3) fp32 += weight gradient. ---- */
int rss_stem_conv_fwd(const void* x, const float* w /*(64,3,3,3)*/, void* y, int B, int H, int W, int in_dtype,
                      float* stat_accum, const float* stat_shift, cudaStream_t stream);
int rss_stem_conv_wgrad(const void* x, const void* dy, float* dw_acc, int B, int H, int W, int in_dtype, cudaStream_t stream);

/* ---- head: hrnet_aux.py:78-81 (Conv2d(480,7,1)); logits are kept at LOW resolution, (pixels, 8) fp32
 *      (class 7 is padding); the x4 UpsamplingBilinear2d is fused into rss_seg_loss_fwd / rss_head_probs ---- */
int rss_head_fwd(const void* x, const float* w /*(7,C)*/, const float* bias, float* logits_lr, int64_t pixels, int C,
                 int dtype, cudaStream_t stream);
int rss_head_bwd(const void* x, const float* dlogits_lr, const float* w, void* dx, float* dw_acc, float* db_acc,
                 int64_t pixels, int C, int dtype, cudaStream_t stream);
/* eval output, hrnet_aux.py:109-110: probs (B,7,h*scale,w*scale) NCHW fp32 = softmax(upsample(logits)); argmax u8 optional */
int rss_head_probs(const float* logits_lr, float* probs, uint8_t* argmax, int B, int h, int w, int scale, cudaStream_t stream);
/* headaux, hrnet_aux.py:86-87,99-101: AdaptiveAvgPool2d(1) + Linear(C,7) (receives no gradient in the reference) */
int rss_headaux_fwd(const void* f0, const float* w /*(7,C)*/, const float* bias, float* colsum_ws /*[B*C]*/, float* scores /*(B,7)*/,
                    int B, int HW, int C, int dtype, cudaStream_t stream);

/* ---- loss: module/CGFL.py:201-227,72-101 + losses/auxloss.py:257-305 (closed form in loss_optim.cu) ----
 * out4 = {loss, grad scale, CE, n_valid}; gdir (B,h,w,8) f32 = un-scaled d loss/d logits_lr */
size_t rss_seg_loss_acc_floats(int B);
int rss_seg_loss_fwd(const float* logits_lr, const int64_t* labels /*(B,h*scale,w*scale)*/, const float* aux_scores /*(B,7)*/,
                     float* acc_ws, float* gdir, float* out4, int B, int h, int w, int scale, int ignore_index, cudaStream_t stream);
int rss_seg_loss_bwd(const float* gdir, const float* out4, const float* upstream /*scalar or NULL*/, float* dlogits_lr,
                     int B, int h, int w, cudaStream_t stream);

/* ---- optimiser step: configs/base/loveda.py:68-99 (clip_grad_norm_(35,L2) + SGD(m=.9, wd=1e-4)) on ONE flat buffer ---- */
int rss_grad_sumsq(const float* grads, int64_t n, float grad_scale, double* sumsq, cudaStream_t stream);
/* momentum_buf must start zeroed (then m = g on the first step, as torch.optim.SGD); lr is a DEVICE scalar (poly schedule
 * written by the host between steps, so a captured CUDA graph of the step stays valid). */
int rss_sgd_step(float* params, float* grads, float* momentum_buf, int64_t n, const double* sumsq, float grad_scale,
                 float max_norm, const float* lr, float momentum, float weight_decay, int zero_grad,
                 void* bf16_shadow /*may be NULL*/, cudaStream_t stream);
/* channels-last (Cout,kh,kw,Cin) bf16 copies of the k>1 convolution weights living in the flat fp32 parameter buffer, all in one
 * launch.  table[e] = {src offset in floats, dst offset in elements, Cin, kh*kw} (int64 x4, device memory); row_start[e] = sum of
 * Cout over the entries before e, row_start[n_entries] = total (int64, device memory); max_row_floats = max Cin*kh*kw. */
int rss_shadow_cl_refresh(const float* params, void* shadow_cl, const int64_t* table, const int64_t* row_start,
                          int n_entries, int max_row_floats, cudaStream_t stream);
/* ---- evaluation side (train.py:17,42-55, eval.py:48-80, module/tta.py:12-24,118-137) ----
 * rss_bilinear_resize: dst = alpha * F.interpolate(src, (H,W), 'bilinear', align_corners=True) + beta * dst on `planes` NCHW fp32
 * planes (beta == 0: dst is not read): the Scale transform of the test-time augmentation and its accumulating inverse.
 * rss_confusion_matrix: cm_acc[truth*K + pred] += 1 over the pixels with truth != ignore_index (PixelMetric.forward), from the
 * uint8 arg-max map of rss_head_probs; cm_acc (K*K uint64) is accumulated into. */
int rss_bilinear_resize(const float* src, float* dst, int planes, int h, int w, int H, int W, float alpha, float beta, cudaStream_t stream);
int rss_confusion_matrix(const uint8_t* pred, const int64_t* truth, unsigned long long* cm_acc, int64_t n, int num_classes,
                         int ignore_index, cudaStream_t stream);

/* dst_e[i] += (float)src_e[i] for a list of tensors in ONE launch (fp32 accumulation of the library's bf16 weight gradients into the
 * flat gradient buffer).  table[e] = {src device pointer (bf16), dst device pointer (f32), numel, Cin, kk}: kk = 0 same element
 * order, kk = kh*kw > 0: src in the library's channels-last (Cout,kh,kw,Cin) order, dst in parameter order (Cout,Cin,kh,kw).
 * chunk_start[e] = index of the first chunk of entry e, chunk_start[n_entries] = total_chunks (device memory); entry e has
 * rss_accum_chunks(numel, Cin, kk) chunks: 4096 elements each, or whole (Cin*kk)-element rows when kk > 0 and a row fits. */
int64_t rss_accum_chunks(int64_t numel, int64_t cin, int64_t kk);
int rss_accum_bf16_list(const int64_t* table, const int64_t* chunk_start, int n_entries, int64_t total_chunks, cudaStream_t stream);
/* transposed bf16 copies for the data-gradient operand of rss_conv_cf: weight e = fp32 (Cout,Cin,kh,kw) at params + table[e][0]
 * -> bf16 [Cin][kh*kw][Cout] at shadow_t + table[e][1]; table[e] = {src offset, dst offset, Cout, Cin, kh*kw} (int64). */
int rss_shadow_t_refresh(const float* params, void* shadow_t, const int64_t* table, int n_entries, cudaStream_t stream);

/* ---- SCD-AAAI2023 DenseEnergyLoss: permutohedral-lattice bilateral filter ----
 * Replaces `void bilateralfilter_batch(float* images, int len_images, float* ins, int len_ins, float* outs, int len_outs, int N,
 * int K, int H, int W, float sigmargb, float sigmaxy)` (SCD-AAAI2023/wrapper/bilateralfilter/bilateralfilter.hpp:12, .cpp:43-55;
 * SWIG typemaps bilateralfilter.i:21-25; caller utils/losses.py:66-70): images (N,3,H,W) de-normalised RGB, ins/outs (N,K,H,W),
 * flat contiguous fp32; outs[n,k] = lattice filter of ins[n,k] guided by (x/sigmaxy, y/sigmaxy, rgb/sigmargb) of image n.
 * Results are bit-identical to the reference built with its own setup.py flags (x86-64 SSE path of permutohedral.cpp).
 *
 * rss_bilateralfilter_batch: DEVICE pointers, stream-ordered, no host synchronisation; `workspace` (>= rss_bilateral_workspace_bytes,
 * 16-byte aligned) is caller-owned scratch; lattice_points (device int[1] or NULL) receives the number of lattice points of the batch.
 * rss_bilateralfilter_batch_host: the reference's argument list verbatim (HOST arrays, outs written in place, lengths checked),
 * synchronous; the only entry point that owns device memory (a process-wide arena grown on demand).
 * rss_dense_energy_gate: the element-wise part of DenseEnergyLossFunction.forward (utils/losses.py:54-64,71-74) fused into one
 * pass over device data: AS[i] *= gate(n,pixel) with gate = 1 where unlabeled, else max(ROI - max_k seg, 0); loss_acc[0] -=
 * sum(seg_roi * AS) / N (double, atomic, caller zeroes it); seg_roi is the ROI-masked segmentation that went into the filter. */
size_t rss_bilateral_workspace_bytes(int N, int K, int H, int W);
int rss_bilateralfilter_batch(const float* images, const float* ins, float* outs, int N, int K, int H, int W, float sigmargb,
                              float sigmaxy, void* workspace, size_t workspace_bytes, int* lattice_points, cudaStream_t stream);
int rss_bilateralfilter_batch_host(const float* images, int len_images, const float* ins, int len_ins, float* outs, int len_outs,
                                   int N, int K, int H, int W, float sigmargb, float sigmaxy);
int rss_dense_energy_gate(const float* seg, const float* rois, const uint8_t* unlabeled, const float* seg_roi, float* AS,
                          double* loss_acc, int N, int K, int H, int W, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RSS_B200_H_ */
