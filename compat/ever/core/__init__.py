from . import registry, logger, config, builder, checkpoint     # noqa: F401
