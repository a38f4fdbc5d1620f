"""import_config('baseline.hrnetw32') -> AttrDict of configs/baseline/hrnetw32.py:config; 'a.b.c value' command-line overrides."""
import ast
import importlib


class AttrDict(dict):
    """dict with attribute access and recursive update (hrnet_aux.py:77 `self.config.neck.in_channels`, :112 `self.config.update(...)`)"""

    def __init__(self, *a, **kw):
        super().__init__()
        self.update(dict(*a, **kw))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def update(self, other=None, **kw):
        other = dict(other or {}, **kw)
        for k, v in other.items():
            if isinstance(v, dict) and not isinstance(v, AttrDict):
                cur = self.get(k)
                if isinstance(cur, AttrDict):
                    cur.update(v)
                else:
                    self[k] = AttrDict(v)
            else:
                self[k] = v


def import_config(config_path, prefix="configs"):
    mod = importlib.import_module("%s.%s" % (prefix, config_path))
    return AttrDict(mod.config)


def apply_overrides(cfg, opts):
    """opts = ['train.num_iters', '1', 'model.params.backbone.pretrained', 'False', ...]"""
    if len(opts) % 2:
        raise ValueError("overrides come in 'dotted.key value' pairs: %r" % (opts,))
    for key, val in zip(opts[0::2], opts[1::2]):
        try:
            val = ast.literal_eval(val)
        except (ValueError, SyntaxError):
            pass
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = val
    return cfg
