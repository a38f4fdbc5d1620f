import numpy as np
import torch


class ToTensor(object):
    """HWC image -> CHW float tensor, mask -> int64 tensor (configs/base/loveda.py:33)"""

    def __call__(self, force_apply=False, **data):
        img = np.ascontiguousarray(data["image"])
        data["image"] = torch.from_numpy(img.transpose(2, 0, 1).astype(np.float32))
        if "mask" in data:
            data["mask"] = torch.from_numpy(np.ascontiguousarray(data["mask"]).astype(np.int64))
        return data
