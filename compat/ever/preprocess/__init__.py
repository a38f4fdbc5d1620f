from . import albu     # noqa: F401
