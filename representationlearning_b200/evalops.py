"""Evaluation-side mirror of the reference: test-time augmentation and the pixel metric on the device.

  tta(model, image, tta_config) / Scale / TestTimeAugmentation   module/tta.py:12-43,118-137
  PixelMetric(num_classes).forward(y_true, y_pred) / summary_all() the `er.metric.PixelMetric` that train.py:17,48-49 and eval.py:48,72 use

The reference resizes with F.interpolate on whatever device the tensors live on and moves every arg-max map to the host for a
numpy confusion matrix.  Here the bilinear (align_corners=True) resize is one kernel with an accumulate epilogue -- the inverse
transform adds each scale's probabilities straight into the running mean -- and the confusion matrix is counted on the device from
the uint8 arg-max map; only the K x K counters are read back, once, in summary_all()."""
import torch

from . import _lib, ops


def bilinear_resize(src, size, out=None, alpha=1.0, beta=0.0):
    """alpha * F.interpolate(src, size, mode='bilinear', align_corners=True) + beta * out  (NCHW fp32)"""
    _lib.require_device()
    src = src.contiguous().float()
    B, C, h, w = src.shape
    H, W = int(size[0]), int(size[1])
    if out is None:
        out = torch.empty(B, C, H, W, device=src.device, dtype=torch.float32)
        beta = 0.0
    assert out.shape == (B, C, H, W) and out.dtype == torch.float32 and out.is_contiguous()
    ops.check(_lib.load().rss_bilinear_resize(src.data_ptr(), out.data_ptr(), B * C, h, w, H, W, float(alpha), float(beta), ops._st()),
              "rss_bilinear_resize")
    return out


class Scale(object):
    """module/tta.py:118-137"""

    def __init__(self, size=None, scale_factor=None):
        self.size, self.scale_factor, self.input_shape = size, scale_factor, None

    def out_size(self, shape):
        if self.size is not None:
            return (self.size, self.size) if isinstance(self.size, int) else tuple(self.size)
        return int(shape[2] * self.scale_factor), int(shape[3] * self.scale_factor)     # F.interpolate floors

    def transform(self, inputs):
        self.input_shape = inputs.shape
        return bilinear_resize(inputs, self.out_size(inputs.shape))

    def inv_transform(self, transformed_inputs, out=None, alpha=1.0, beta=0.0):
        return bilinear_resize(transformed_inputs, (self.input_shape[2], self.input_shape[3]), out, alpha, beta)


@torch.no_grad()
def tta(model, image, tta_config):
    """mean over the transforms of inv_transform(model(transform(image)))   (module/tta.py:12-24)"""
    n = len(tta_config)
    out = None
    for i, t in enumerate(tta_config):
        probs = model(t.transform(image))
        if isinstance(t, Scale):
            out = t.inv_transform(probs, out, alpha=1.0 / n, beta=0.0 if out is None else 1.0)
        else:                                     # index transforms (flips, rot90) stay torch views
            back = t.inv_transform(probs).float() / n
            out = back.contiguous() if out is None else out.add_(back)
    return out


class TestTimeAugmentation(torch.nn.Module):
    """module/tta.py:27-43"""
    __test__ = False

    def __init__(self, module, tta_config):
        super().__init__()
        self.module, self.tta_config = module, tta_config

    @torch.no_grad()
    def forward(self, image):
        return tta(self.module, image, self.tta_config)


class PixelMetric(object):
    """confusion matrix over the valid pixels -> per-class IoU / F1, OA, mIoU; counters live on the device"""

    def __init__(self, num_classes, logdir=None, logger=None, ignore_index=-1, device="cuda"):
        self.num_classes, self.logdir, self.logger, self.ignore_index = int(num_classes), logdir, logger, ignore_index
        self.cm = torch.zeros(self.num_classes * self.num_classes, dtype=torch.int64, device=device)

    def forward(self, y_true, y_pred):
        """y_true int64 labels, y_pred uint8/int64 class indices (same shape; ignore_index pixels of y_true are skipped -- the
        reference filters them on the host with `valid_inds = y_true != -1`, train.py:47-49)"""
        _lib.require_device()
        t = y_true.to(self.cm.device, torch.int64).contiguous().reshape(-1)
        p = y_pred.to(self.cm.device).to(torch.uint8).contiguous().reshape(-1)
        assert t.numel() == p.numel()
        ops.check(_lib.load().rss_confusion_matrix(p.data_ptr(), t.data_ptr(), self.cm.data_ptr(), t.numel(), self.num_classes,
                                                   self.ignore_index, ops._st()), "rss_confusion_matrix")

    def update_from_probs(self, logits_lr_or_probs, y_true, model=None):
        """eval loop helper: arg-max on the device (probabilities (B,K,H,W))"""
        self.forward(y_true, logits_lr_or_probs.argmax(dim=1))

    def summary_all(self):
        cm = self.cm.view(self.num_classes, self.num_classes).double().cpu()
        tp = cm.diag()
        iou = tp / torch.clamp(cm.sum(0) + cm.sum(1) - tp, min=1)
        f1 = 2 * tp / torch.clamp(cm.sum(0) + cm.sum(1), min=1)
        out = dict(iou=iou.numpy(), f1=f1.numpy(), miou=float(iou.mean()), oa=float(tp.sum() / max(float(cm.sum()), 1.0)), cm=cm.numpy())
        msg = "mIoU %.4f  OA %.4f  IoU %s" % (out["miou"], out["oa"], [round(float(v), 4) for v in iou])
        (self.logger.info if self.logger is not None else print)(msg)
        return out
