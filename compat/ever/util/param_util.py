def count_model_parameters(module, logger=None):
    n = sum(p.numel() for p in module.parameters())
    (logger.info if logger is not None else print)("# parameters: %.2f M" % (n / 1e6))
    return n
