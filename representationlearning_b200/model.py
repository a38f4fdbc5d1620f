"""`RSSFormer` — the registered model of RSSFormer-TIP2023/module/baseline/hrnet_aux.py:70-134,
rebuilt on the sm_100a kernels.  Same registry name, constructor (`HRNetFusion(config)`), forward
contract (`model(img)` -> (B,7,H,W) softmax in eval; `model(img, {'cls': labels})` -> {'fc_loss': scalar}
in train) and state_dict keys as the reference.
"""
import torch
import torch.nn as nn

from . import ops
from .hrnet import hrnetv2_w32
from .modules import SimpleFusion8


class AttrDict(dict):
    """Attribute-style config with recursive update (the subset of `ever`'s config object the model uses:
    hrnet_aux.py:77 `self.config.neck.in_channels`, :112 `self.config.update(dict(...))`)."""

    def __init__(self, *a, **kw):
        super().__init__()
        self.update(dict(*a, **kw))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def update(self, other=None, **kw):
        for k, v in dict(other or {}, **kw).items():
            if isinstance(v, dict):
                cur = self.get(k)
                if isinstance(cur, AttrDict):
                    cur.update(v)
                else:
                    self[k] = AttrDict(v)
            else:
                self[k] = v


class Registry(dict):
    """`ever.registry.MODEL`-style registry (hrnet_aux.py:70, hrnet_encoder.py:14-17,32)."""

    def register(self, name=None, obj=None):
        if obj is not None:
            self[name] = obj
            return obj

        def deco(o):
            self[name or o.__name__] = o
            return o
        return deco


MODEL = Registry()
MODEL.register("hrnetv2_w32", hrnetv2_w32)

# configs/baseline/hrnetw32.py:7-33 restated (pretrained off: no network/checkpoint offline)
RSSFORMER_CONFIG = dict(
    backbone=dict(hrnet_type="hrnetv2_w32", pretrained=False, norm_eval=False, frozen_stages=-1, with_cp=False, with_gc=False),
    neck=dict(in_channels=480), classes=7, head=dict(in_channels=480, upsample_scale=4.0),
    loss=dict(ignore_index=-1, ce=dict()),
)


class ERModule(nn.Module):
    def __init__(self, config=None):
        super().__init__()
        self._cfg = AttrDict()
        self.set_default_config()
        self._cfg.update(dict(config or {}))

    @property
    def config(self):
        return self._cfg

    def set_default_config(self):
        pass


@MODEL.register("HRNetEncoder")
class HRNetEncoder(ERModule):
    """base_hrnet/hrnet_encoder.py:28-41"""

    def __init__(self, config=None):
        super().__init__(config)
        self.hrnet = MODEL[self.config.hrnet_type](pretrained=self.config.pretrained, weight_path=self.config.weight_path,
                                                   norm_eval=self.config.norm_eval, frozen_stages=self.config.frozen_stages)

    def forward(self, x):
        return self.hrnet(x)

    def set_default_config(self):
        self.config.update(dict(hrnet_type="hrnetv2_w18", pretrained=False, weight_path=None, norm_eval=False,
                                frozen_stages=-1, with_cp=False))

    def output_channels(self):
        if self.config.hrnet_type == "hrnetv2_w32":
            return 32, 64, 128, 256
        raise NotImplementedError("{} is not implemented.".format(self.config.hrnet_type))


class _Head(nn.Sequential):
    """nn.Sequential(Conv2d(480,7,1), UpsamplingBilinear2d(x4)) of hrnet_aux.py:78-81.  Kept as a Sequential for the
    state_dict keys (`head.0.*`); the conv runs through rss_head_fwd and the up-sampling is fused downstream."""

    def logits_lr(self, x):
        return ops.HeadConv.apply(x, self[0].weight, self[0].bias)

    def forward(self, x):
        """full-resolution logits (B,7,H,W) — only for callers that want them materialised."""
        lr = self.logits_lr(x)[..., :7].permute(0, 3, 1, 2)
        return nn.functional.interpolate(lr, scale_factor=float(self[1].scale_factor), mode="bilinear", align_corners=True)


class SegmentationLossaux(nn.Module):
    """module/CGFL.py:192-227 (the `ce` branch, the only one the RSSFormer config enables: hrnetw32.py:24-26)."""

    def __init__(self, loss_config):
        super().__init__()
        self.loss_config = loss_config
        for k in ("fcloss", "bceloss", "tverloss", "diceloss"):
            if k in loss_config:
                raise NotImplementedError("loss term '%s' is not enabled in the RSSFormer config" % k)

    def forward(self, logits_lr, y_true, aux_scores, scale=4):
        loss_dict = dict()
        if "ce" in self.loss_config:
            loss_dict["fc_loss"] = ops.SegLoss.apply(logits_lr, y_true, aux_scores, scale, self.loss_config.get("ignore_index", -1))
        return loss_dict


@MODEL.register("RSSFormer")
class HRNetFusion(ERModule):
    def __init__(self, config=None):
        super().__init__(config)
        self.backbone = HRNetEncoder(self.config.backbone)
        self.neck = SimpleFusion8(self.config.neck.in_channels)
        self.head = _Head(nn.Conv2d(self.config.head.in_channels, self.config.classes, 1),
                          nn.UpsamplingBilinear2d(scale_factor=self.config.head.upsample_scale))
        self.loss = SegmentationLossaux(self.config.loss)
        self.headaux = nn.Sequential(nn.Linear(32, 7))
        self.compute_dtype = torch.bfloat16          # activation dtype inside the model (fp32 for strict-parity runs)

    def forward(self, x, y=None):
        # the image batch goes to the backbone as it came in ((B,3,H,W), usually planar fp32): the stem casts it -- inside its first
        # convolution kernel when that applies (hrnet.HighResolutionNet._stem1)
        self.backbone.hrnet.stem_dtype = self.compute_dtype
        feats = self.backbone(x)
        fused, f0 = self.neck(feats)
        aux = ops.headaux(f0, self.headaux[0].weight, self.headaux[0].bias)
        logits_lr = self.head.logits_lr(fused)
        scale = int(self.config.head.upsample_scale)
        if self.training:
            return self.loss(logits_lr, y["cls"].long(), aux, scale)
        return ops.head_probs(logits_lr, scale)

    def set_default_config(self):
        self.config.update(dict(
            backbone=dict(hrnet_type="hrnetv2_w48", pretrained=False, norm_eval=False, frozen_stages=-1, with_cp=False, with_gc=False),
            neck=dict(in_channels=720), classes=7, head=dict(in_channels=720, upsample_scale=4.0), loss=dict(ce=dict())))


def build_rssformer(config=None, compute_dtype=torch.bfloat16, device="cuda"):
    m = HRNetFusion(config or RSSFORMER_CONFIG)
    m.compute_dtype = compute_dtype
    return m.to(device)
