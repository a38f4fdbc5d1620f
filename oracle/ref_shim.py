"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Import shim that lets THIS container import the *unmodified* reference
(`/root/reference/RSSFormer-TIP2023`) so that `oracle/gen_golden.py` can
(1) pin the CPU restatement in `oracle/rssformer_ref.py` against the real code
and (2) write the golden vectors committed under `tests/golden/`.

The reference delegates registry/config plumbing to the un-vendored, un-pinned
`ever` package (RSSFormer-TIP2023/train.py:1,79-80) and uses two helpers of
`timm` (modules/multihead_isa_attention.py:12).  Neither is installed here, so
minimal stand-ins are registered in `sys.modules` *before* the reference is
imported.  Nothing in here does arithmetic: the numerics that come out of
`load_reference()` are 100 % the reference's own code on torch CPU.

`/root/reference` does not exist on the GPU box.  There the same unmodified sources are imported from the
travelling archive `oracle/_ref/rssformer_reference.zip` (zipimport; built by `oracle/build_ref.py`, git-ignored).
This module must only be used from `oracle/gen_golden.py`, `bench.py`'s reference / baseline legs and tests that
skip when neither the tree nor the archive is present.
"""
import os
import sys
import types
import logging

REFERENCE_ROOT = os.environ.get("RSS_REFERENCE_ROOT", "/root/reference/RSSFormer-TIP2023")
REFERENCE_ZIP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "rssformer_reference.zip")


def reference_tree_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "module", "baseline"))


def reference_available():
    """the unmodified reference can be imported: from its tree (authoring container) or from the travelling archive"""
    return reference_tree_available() or os.path.exists(REFERENCE_ZIP)


def reference_path():
    return REFERENCE_ROOT if reference_tree_available() else REFERENCE_ZIP


class _AttrDict(dict):
    """dict with attribute access and recursive update (what `ever`'s config object offers:
    hrnet_aux.py:77 uses `self.config.neck.in_channels`, :112 `self.config.update(dict(...))`)."""

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
        other = dict(other or {}, **kw)
        for k, v in other.items():
            if isinstance(v, dict):
                cur = self.get(k)
                if isinstance(cur, _AttrDict):
                    cur.update(v)
                else:
                    self[k] = _AttrDict(v)
            else:
                self[k] = v


class _Registry(dict):
    def register(self, name=None, obj=None):
        if obj is not None:
            self[name] = obj
            return obj

        def deco(o):
            self[name or o.__name__] = o
            return o
        return deco


def _install_fakes():
    import torch.nn as nn
    if "ever" in sys.modules and getattr(sys.modules["ever"], "_rss_shim", False):
        return
    ever = types.ModuleType("ever")
    ever._rss_shim = True
    core = types.ModuleType("ever.core")
    registry = types.ModuleType("ever.core.registry")
    registry.MODEL = _Registry()
    registry.DATALOADER = _Registry()
    registry.register_all = lambda: None
    logger = types.ModuleType("ever.core.logger")
    logger.get_logger = lambda *a, **k: logging.getLogger("ever-shim")
    interface = types.ModuleType("ever.interface")

    class ConfigurableMixin(object):
        def __init__(self, config=None):
            self._cfg = _AttrDict()
            self.set_default_config()
            self._cfg.update(dict(config or {}))

        @property
        def config(self):
            return self._cfg

        def set_default_config(self):
            pass

    class ERModule(nn.Module, ConfigurableMixin):
        def __init__(self, config=None):
            nn.Module.__init__(self)
            ConfigurableMixin.__init__(self, config)

    interface.ERModule = ERModule
    interface.ConfigurableMixin = ConfigurableMixin
    core.registry = registry
    core.logger = logger
    ever.core = core
    ever.registry = registry
    ever.interface = interface
    ever.ERModule = ERModule
    sys.modules.update({
        "ever": ever, "ever.core": core, "ever.core.registry": registry,
        "ever.core.logger": logger, "ever.interface": interface,
    })

    timm = types.ModuleType("timm")
    tm = types.ModuleType("timm.models")
    tl = types.ModuleType("timm.models.layers")
    tl.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
    tl.trunc_normal_ = nn.init.trunc_normal_
    timm.models = tm
    tm.layers = tl
    sys.modules.setdefault("timm", timm)
    sys.modules.setdefault("timm.models", tm)
    sys.modules.setdefault("timm.models.layers", tl)


RSSFORMER_PARAMS = dict(  # restated from configs/baseline/hrnetw32.py:7-33 (pretrained forced off: no network)
    backbone=dict(hrnet_type="hrnetv2_w32", pretrained=False, norm_eval=False,
                  frozen_stages=-1, with_cp=False, with_gc=False),
    neck=dict(in_channels=480),
    classes=7,
    head=dict(in_channels=480, upsample_scale=4.0),
    loss=dict(ignore_index=-1, ce=dict()),
)


def load_reference():
    """Returns the reference's python modules (hrnet_aux, MTFM, ...) imported from REFERENCE_ROOT."""
    if not reference_available():
        raise RuntimeError("reference not present at %s or %s" % (REFERENCE_ROOT, REFERENCE_ZIP))
    _install_fakes()
    if reference_path() not in sys.path:
        sys.path.insert(0, reference_path())
    import importlib
    ns = types.SimpleNamespace()
    ns.hrnet_aux = importlib.import_module("module.baseline.hrnet_aux")
    ns.MTFM = importlib.import_module("module.baseline.base_hrnet.modules.MTFM")
    ns.pool = importlib.import_module("module.baseline.base_hrnet.modules.multihead_isa_pool_attention")
    ns.DAL = importlib.import_module("module.baseline.base_hrnet.modules.DAL")
    ns.ffn = importlib.import_module("module.baseline.base_hrnet.modules.ffn_block")
    ns.CGFL = importlib.import_module("module.CGFL")
    ns.hrnet = importlib.import_module("module.baseline.base_hrnet._hrnet_rssformer")
    return ns


def build_reference_model(seed=2333):
    """HRNetFusion exactly as train.py would build it (seed: train.py:78), pretrained off."""
    import torch
    ns = load_reference()
    torch.manual_seed(seed)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        model = ns.hrnet_aux.HRNetFusion(RSSFORMER_PARAMS)
    return model
