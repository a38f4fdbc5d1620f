from . import param_util     # noqa: F401
