"""Minimal stand-in for the albumentations API that configs/base/loveda.py and data/loveda.py use (see compat/README.md):
dict-in/dict-out transforms on numpy HWC images + HW masks.  Geometric warps beyond flips / 90-degree rotations / crops are
identity (ShiftScaleRotate, blurs): they only matter for accuracy, which no test here measures."""
import random

import numpy as np


class _T(object):
    def __init__(self, *a, p=0.5, always_apply=False, **kw):
        self.p = 1.0 if (always_apply or (a and a[0] is True)) else p

    def apply(self, image, mask):
        return image, mask

    def __call__(self, force_apply=False, **data):
        if force_apply or random.random() < self.p:
            img, m = self.apply(data["image"], data.get("mask"))
            data["image"] = img
            if m is not None:
                data["mask"] = m
        return data


class Compose(object):
    def __init__(self, transforms, **kw):
        self.transforms = list(transforms)

    def __call__(self, force_apply=False, **data):
        for t in self.transforms:
            data = t(**data)
        return data


class OneOf(object):
    def __init__(self, transforms, p=0.5):
        self.transforms, self.p = list(transforms), p

    def __call__(self, force_apply=False, **data):
        if self.transforms and random.random() < self.p:
            return random.choice(self.transforms)(force_apply=True, **data)
        return data


class HorizontalFlip(_T):
    def apply(self, image, mask):
        return image[:, ::-1], None if mask is None else mask[:, ::-1]


class VerticalFlip(_T):
    def apply(self, image, mask):
        return image[::-1], None if mask is None else mask[::-1]


class RandomRotate90(_T):
    def apply(self, image, mask):
        k = random.randint(0, 3)
        return np.rot90(image, k), None if mask is None else np.rot90(mask, k)


class RandomCrop(_T):
    def __init__(self, height, width, p=1.0, **kw):
        self.h, self.w, self.p = height, width, p

    def apply(self, image, mask):
        H, W = image.shape[:2]
        y, x = random.randint(0, max(H - self.h, 0)), random.randint(0, max(W - self.w, 0))
        return image[y:y + self.h, x:x + self.w], None if mask is None else mask[y:y + self.h, x:x + self.w]


class Normalize(_T):
    def __init__(self, mean=(), std=(), max_pixel_value=255.0, always_apply=True, p=1.0):
        self.mean = np.asarray(mean, dtype=np.float32) * max_pixel_value
        self.std = np.asarray(std, dtype=np.float32) * max_pixel_value
        self.p = 1.0

    def apply(self, image, mask):
        img = image.astype(np.float32)
        if self.mean.size:
            img = (img - self.mean) / self.std
        return img, mask


class _Identity(_T):
    pass


ShiftScaleRotate = MotionBlur = MedianBlur = Blur = RandomScale = RandomBrightnessContrast = _Identity
