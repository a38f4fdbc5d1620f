"""name -> class registries and register_all() (imports the project modules whose decorators fill them)."""
import importlib
import os


class Registry(dict):
    def register(self, name=None, obj=None):
        if obj is not None:
            self[name] = obj
            return obj

        def deco(o):
            self[name or o.__name__] = o
            return o
        return deco


MODEL = Registry()
DATALOADER = Registry()
OPT = Registry()
LR = Registry()


def register_all():
    """the reference calls this at import time (train.py:11).  Modules to import come from EVER_REGISTER (comma separated;
    default: the RSSFormer model and the LoveDA loader).  RSS_IMPL=b200 re-binds 'RSSFormer' to this repo's model."""
    mods = os.environ.get("EVER_REGISTER", "module.baseline.hrnet_aux,data.loveda").split(",")
    for m in mods:
        if m:
            importlib.import_module(m)
    if os.environ.get("RSS_IMPL", "reference") == "b200":
        import representationlearning_b200 as P
        MODEL["RSSFormer"] = P.HRNetFusion
