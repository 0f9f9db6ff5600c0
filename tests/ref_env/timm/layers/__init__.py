"""timm.layers names imported by network/model_parts.py:34 (DropPath is unused there)."""
import torch.nn as nn
from torch.nn.init import trunc_normal_  # noqa: F401


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class DropPath(nn.Identity):
    def __init__(self, drop_prob=0.0, *a, **kw):
        super().__init__()
