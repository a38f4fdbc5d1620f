import torch.nn as nn

from ..core.config import AttrDict


class ConfigurableMixin(object):
    def __init__(self, config=None):
        self._cfg = AttrDict()
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
