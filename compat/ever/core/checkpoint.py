import torch


def remove_module_prefix(state_dict):
    sd = state_dict.get("model", state_dict) if isinstance(state_dict, dict) else state_dict
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def load_model_state_dict_from_ckpt(path):
    return remove_module_prefix(torch.load(path, map_location="cpu"))
