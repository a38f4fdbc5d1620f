"""ctypes binding of the C-ABI library `librss_b200.so` (declared in include/rss_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing, or the device is not an
sm_100 part, every entry point raises.  This is the same stub a maintainer of the reference would
add to bind the library from `RSSFormer-TIP2023/module/...`.
"""
import ctypes
import os
from ctypes import c_int, c_int64, c_float, c_size_t, c_void_p, POINTER, Structure

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librss_b200.so")

RSS_F32, RSS_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
_STATUS = {0: "RSS_OK", -1: "RSS_ERR_SHAPE", -2: "RSS_ERR_DTYPE", -3: "RSS_ERR_CUDA", -4: "RSS_ERR_WORKSPACE", -5: "RSS_ERR_ARCH"}


class RssError(RuntimeError):
    pass


class AttnParams(Structure):
    _fields_ = [(n, c_void_p) for n in ("ln_w", "ln_b", "sa1_w", "sa2_w", "lvl_w", "lvl_b",
                                        "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b")] + \
               [("ln_eps", c_float), ("C", c_int), ("num_heads", c_int), ("window", c_int)]


class AttnGrads(Structure):
    _fields_ = [(n, c_void_p) for n in ("ln_w", "ln_b", "sa1_w", "sa2_w", "lvl_w", "lvl_b",
                                        "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b")]


class ConvCfEpilogue(Structure):
    """RssConvCfEpilogue of include/rss_b200.h"""
    _fields_ = [("mode", c_int), ("add", c_void_p), ("accum", c_void_p), ("ticket", c_void_p), ("gamma", c_void_p), ("beta", c_void_p),
                ("running_mean", c_void_p), ("running_var", c_void_p), ("momentum", c_float), ("eps", c_float),
                ("mean_out", c_void_p), ("invstd_out", c_void_p), ("scale_out", c_void_p), ("shift_out", c_void_p),
                ("bn_z", c_void_p), ("bn_out", c_void_p), ("bn_mean", c_void_p), ("bn_invstd", c_void_p), ("bn_scale", c_void_p),
                ("bn_shift", c_void_p), ("bn_relu", c_int), ("sums_out", c_void_p)]


CF_PLAIN, CF_STATS, CF_BNRED = 0, 1, 2
P = c_void_p
# name -> (restype, argtypes); must list every symbol include/rss_b200.h declares (tests check this)
SIGNATURES = {
    "rss_version": (c_int, []),
    "rss_last_cuda_error": (c_int, []),
    "rss_check_device": (c_int, []),
    "rss_layernorm_fwd": (c_int, [P, P, P, P, P, P, c_float, c_int64, c_int, c_int, P]),
    "rss_layernorm_bwd": (c_int, [P, P, P, P, P, P, P, P, P, c_int64, c_int, c_int, P]),
    "rss_attn_fwd": (c_int, [P, P, POINTER(AttnParams), c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P]),
    "rss_spatial_attention_fwd": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "rss_attn_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "rss_attn_bwd": (c_int, [P, P, P, POINTER(AttnParams), c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, c_size_t,
                             P, P, POINTER(AttnGrads), P]),
    "rss_bn_stats_nparts": (c_int, [c_int64, c_int]),
    "rss_bn_stats": (c_int, [P, P, P, c_int64, c_int, c_int, P]),
    "rss_bn_stats_fused": (c_int, [P, P, P, c_int64, c_int, c_int, P, P, P, P, c_float, c_float, P, P, P, P, P, P]),
    "rss_bn_combine": (c_int, [P, P, c_int, c_int, P, P, P]),
    "rss_bn_finalize": (c_int, [P, P, P, P, P, P, c_float, c_float, c_int, P, P, P, P, P, P]),
    "rss_bn_eval_affine": (c_int, [P, P, P, P, c_float, c_int, P, P, P, P, P]),
    "rss_bn_act_fwd": (c_int, [P, P, P, P, P, c_int64, c_int, c_int, c_int, P]),
    "rss_bn_bwd_reduce": (c_int, [P, P, P, P, P, P, P, P, c_int64, c_int, c_int, c_int, P]),
    "rss_bn_bwd_reduce_ws": (c_int, [P, P, P, P, P, P, P, P, P, P, P, c_int64, c_int, c_int, c_int, P]),
    "rss_bn_stats_raw": (c_int, [P, P, c_int64, c_int, c_int, P, P, P]),
    "rss_bn_act_fwd_raw": (c_int, [P, P, P, P, P, c_int64, c_int, c_int, c_int, P, P, P, P, c_float, c_float, P, P, P, P, P, P]),
    "rss_bn_bwd_apply_raw": (c_int, [P, P, P, P, P, P, P, P, P, c_float, P, P, c_int64, c_int, c_int, c_int, P, P, P, P]),
    "rss_bn_bwd_apply_dz": (c_int, [P, P, P, P, P, P, c_float, P, c_int64, c_int, c_int, P, P, P, P]),
    "rss_bn_bwd_apply": (c_int, [P, P, P, P, P, P, P, P, c_float, P, P, c_int64, c_int, c_int, c_int, P, P, P, P]),
    "rss_sync_exchange_bytes": (c_size_t, [c_int]),
    "rss_sync_allreduce_small": (c_int, [P, c_int64, c_int, c_int, P, P, c_int, P, P, P]),
    "rss_sync_bn_finalize": (c_int, [P, c_int64, c_int, c_int, P, P, c_int, c_int64, P, P, P, P, c_float, c_float, P, P, P, P, P, P]),
    "rss_fuse_sum_fwd": (c_int, [POINTER(c_void_p), POINTER(c_int), c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "rss_fuse_sum_bwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "rss_conv_igemm_supported": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "rss_conv_packed_bytes": (c_size_t, [c_int, POINTER(c_int), c_int, c_int]),
    "rss_conv_pack_weights": (c_int, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int), POINTER(c_int), c_int, c_int, c_int, c_int,
                                      P, P, POINTER(c_int), POINTER(c_int), POINTER(c_int), P]),
    "rss_conv_igemm": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), P]),
    "rss_conv_igemm_stats": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), P, P, P, P]),
    "rss_conv_cf_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "rss_conv_cf": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), c_int, c_int,
                            P, P, c_int, POINTER(ConvCfEpilogue), P]),
    "rss_conv_wgrad_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "rss_conv_wgrad": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "rss_conv_wgrad_tc_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "rss_conv_wgrad_tc": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P]),
    "rss_neck_gather_fwd": (c_int, [P, P, P, P, P, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int), c_int, P]),
    "rss_neck_gather_bwd": (c_int, [P, P, P, P, P, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int), c_int, P]),
    "rss_stem_conv_fwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P, P, P]),
    "rss_stem_conv_wgrad": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P]),
    "rss_head_fwd": (c_int, [P, P, P, P, c_int64, c_int, c_int, P]),
    "rss_head_bwd": (c_int, [P, P, P, P, P, P, c_int64, c_int, c_int, P]),
    "rss_head_probs": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P]),
    "rss_headaux_fwd": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "rss_seg_loss_acc_floats": (c_size_t, [c_int]),
    "rss_seg_loss_fwd": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "rss_seg_loss_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, P]),
    "rss_grad_sumsq": (c_int, [P, c_int64, c_float, P, P]),
    "rss_sgd_step": (c_int, [P, P, P, c_int64, P, c_float, c_float, P, c_float, c_float, c_int, P, P]),
    "rss_shadow_cl_refresh": (c_int, [P, P, P, P, c_int, c_int, P]),
    "rss_shadow_t_refresh": (c_int, [P, P, P, c_int, P]),
    "rss_accum_bf16_list": (c_int, [P, P, c_int, c_int64, P]),
    "rss_accum_chunks": (c_int64, [c_int64, c_int64, c_int64]),
    "rss_bilinear_resize": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_float, c_float, P]),
    "rss_confusion_matrix": (c_int, [P, P, P, c_int64, c_int, c_int, P]),
    "rss_bilateral_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "rss_bilateralfilter_batch": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_float, c_float, P, c_size_t, P, P]),
    "rss_bilateralfilter_batch_host": (c_int, [P, c_int, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_float, c_float]),
    "rss_dense_energy_gate": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
}

_lib = None


def load():
    """dlopen the library and attach prototypes. Raises RssError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RssError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(no CPU/PyTorch fallback exists for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        lib = load()
        raise RssError("%s failed: %s (cuda error %d)" % (what, _STATUS.get(rc, rc), lib.rss_last_cuda_error()))


_device_ok = False
_device_index = None


def require_device():
    """The product path refuses to run anywhere but on an sm_100 GPU -- and on ONE GPU per process (the launch model of bench.py /
    torchrun): the kernels cache per-device launch attributes (opt-in shared memory size, SM count) for the first device used."""
    global _device_ok, _device_index
    import torch
    if _device_ok:
        if torch.cuda.current_device() != _device_index:
            raise RssError("librss_b200 drives one GPU per process (first used cuda:%d, now cuda:%d): launch one process per GPU"
                           % (_device_index, torch.cuda.current_device()))
        return
    if not torch.cuda.is_available():
        raise RssError("no CUDA device: representationlearning_b200 has no CPU fallback")
    check(load().rss_check_device(), "rss_check_device")
    _device_index = torch.cuda.current_device()
    _device_ok = True
