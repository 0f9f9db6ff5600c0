"""`MSUNet(config, img_size, num_classes)` — same constructor, forward, freeze and pretrained-weight
API as the reference's network/MSUNet.py:16-229, running on the sm_100a kernels."""
from __future__ import annotations

import logging
import os

import torch
import torch.nn as nn

from .model_parts import MSUNetSys

logger = logging.getLogger(__name__)

# checkpoint-prefix -> MS-UNet prefix (network/MSUNet.py:83-119 for SegFace "backbone.0.*",
# :167-205 for torchvision ImageNet "features.*"); stage 5 holds the depth-18 blocks.
_STAGE_MAP = {"0.0": "patch_embed.proj", "0.2": "patch_embed.norm", "2": "layers.0.downsample",
              "4": "layers.1.downsample", "6": "layers.2.downsample"}
_BLOCK_STAGES = {"1": 0, "3": 1, "5": 2, "7": 3}


def _remap(key: str, root: str):
    """`root` + '0.0.weight' style key -> MS-UNet encoder key, or None if the key is not an encoder entry."""
    rest = key[len(root):]
    parts = rest.split(".")
    for n in (2, 1):
        head = ".".join(parts[:n])
        if head in _STAGE_MAP and not (n == 1 and head == "0"):
            return _STAGE_MAP[head] + "." + ".".join(parts[n:])
    if parts[0] in _BLOCK_STAGES and len(parts) > 2 and parts[1].isdigit():
        return f"layers.{_BLOCK_STAGES[parts[0]]}.blocks.{parts[1]}." + ".".join(parts[2:])
    return None


class MSUNet(nn.Module):
    def __init__(self, config, img_size=1024, num_classes=1, zero_head=False, vis=False):
        super().__init__()
        self.num_classes = num_classes
        self.zero_head = zero_head
        self.config = config
        sw = config.MODEL.SWIN
        self.ms_unet = MSUNetSys(img_size=img_size, patch_size=sw.PATCH_SIZE, in_chans=sw.IN_CHANS,
                                 num_classes=self.num_classes, embed_dim=sw.EMBED_DIM, depths=sw.DEPTHS,
                                 num_heads=sw.NUM_HEADS, window_size=sw.WINDOW_SIZE, mlp_ratio=sw.MLP_RATIO,
                                 qkv_bias=sw.QKV_BIAS, qk_scale=None, drop_rate=config.MODEL.DROP_RATE,
                                 drop_path_rate=config.MODEL.DROP_PATH_RATE, ape=sw.APE, patch_norm=sw.PATCH_NORM,
                                 use_checkpoint=config.TRAIN.USE_CHECKPOINT,
                                 attn_drop_rate=config.MODEL.ATTN_DROP_RATE)

    def forward(self, x):
        if x.size()[1] != 3:
            msg = f"Expected 3 channels, but got {x.size(1)}"
            logger.error(msg)
            raise ValueError(msg)
        return self.ms_unet(x)

    def freeze_encoder(self, freeze):
        self.ms_unet.freeze_encoder(freeze)

    def unfreeze_encoder(self, layer_num):
        self.ms_unet.unfreeze_encoder(layer_num)

    def _load_encoder(self, path, log, what, root, unwrap=None, skip_prefix=None):
        if not path or not os.path.exists(path):
            log.error(f"No {what} pretrain found at: {path}")
            return
        device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        ckpt = torch.load(path, map_location=device)
        if unwrap is not None:
            if unwrap not in ckpt:
                msg = f"'{unwrap}' not found in checkpoint: {path}"
                log.error(msg)
                raise KeyError(msg)
            ckpt = ckpt[unwrap]
        new_sd, seen = {}, False
        for k, v in ckpt.items():
            if not k.startswith(root.split(".")[0]):
                continue
            seen = True
            if skip_prefix is not None and k.startswith(skip_prefix):
                continue
            nk = _remap(k, root) if k.startswith(root) else None
            if nk is None:
                msg = f"Key {k} not found in dictionary!!"
                log.error(msg)
                raise ValueError(msg)
            new_sd[nk] = v
        if not seen:
            msg = "No new keys from backbone!!"
            log.error(msg)
            raise ValueError(msg)
        own = self.ms_unet.state_dict()
        for k, v in new_sd.items():
            if k in own and v.shape != own[k].shape:
                msg = f"Key {k} does not match the dictionary of MSUNet!"
                log.error(msg)
                raise ValueError(msg)
        self.ms_unet.load_state_dict(new_sd, strict=False)
        log.info(f"End of the {what} pretrained copying process")

    def load_segface_weight(self, config, logging):
        self._load_encoder(config.MODEL.PRETRAIN_SEGFACE, logging, "segface", "backbone.0.",
                           unwrap="state_dict_backbone", skip_prefix="backbone.1.")

    def load_IMAGENET1K_weight(self, config, logging):
        self._load_encoder(config.MODEL.PRETRAIN_IMAGENET1K, logging, "IMAGENET1K", "features.")
