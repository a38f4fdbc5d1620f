from . import registry


def make_model(cfg):
    return registry.MODEL[cfg["type"]](cfg["params"])


def make_dataloader(cfg):
    return registry.DATALOADER[cfg["type"]](cfg["params"])
