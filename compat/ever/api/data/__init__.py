from . import distributed                         # noqa: F401
from .distributed import CrossValSamplerGenerator  # noqa: F401
