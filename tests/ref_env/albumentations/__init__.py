"""albumentations stand-in (dataset/dataset.py:24-33): the colour augmentations are host-side data preparation outside the hot
path; here every transform is accepted and the pipeline applies a small deterministic brightness jitter so that the augmented
branch of RandomGenerator.__call__ (:46-48) still returns a uint8 HWC image."""
import random

import numpy as np


class _T:
    def __init__(self, *a, p=1.0, **kw):
        self.p = p

    def apply(self, img):
        return img


class ToGray(_T): pass
class HueSaturationValue(_T): pass
class RandomGamma(_T): pass
class GaussianBlur(_T): pass


class RandomBrightnessContrast(_T):
    def __init__(self, brightness_limit=0.2, contrast_limit=0.2, p=0.5, **kw):
        super().__init__(p=p)
        self.b = brightness_limit

    def apply(self, img):
        d = random.uniform(-self.b, self.b) * 255.0
        return np.clip(img.astype(np.float32) + d, 0, 255).astype(np.uint8)


class OneOf(_T):
    def __init__(self, transforms, p=0.5):
        super().__init__(p=p)
        self.transforms = transforms

    def apply(self, img):
        return random.choice(self.transforms).apply(img)


class Compose:
    def __init__(self, transforms, **kw):
        self.transforms = transforms

    def __call__(self, image=None, **kw):
        for t in self.transforms:
            if random.random() < t.p:
                image = t.apply(image)
        return {"image": image, **kw}
