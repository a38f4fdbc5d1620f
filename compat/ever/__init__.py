"""Minimal work-alike of the `ever` package for RSSFormer-TIP2023's train.py / eval.py (see compat/README.md)."""
import numpy as _np

if not hasattr(_np, "long"):          # data/loveda.py:84 uses np.long (removed in numpy 1.24)
    _np.long = _np.int64

from .core import registry, logger as _logger_mod            # noqa: E402
from .core.logger import get_logger, info                     # noqa: E402
from .interface import ERModule, ConfigurableMixin            # noqa: E402
from .interface.transform_base import Transform, MultiTransform   # noqa: E402
from . import metric, preprocess, trainer, api, util          # noqa: E402

__all__ = ["registry", "ERModule", "ConfigurableMixin", "MultiTransform", "Transform", "metric", "preprocess", "trainer", "api",
           "util", "get_logger", "info"]
