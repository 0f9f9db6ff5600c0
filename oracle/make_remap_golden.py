"""Generate tests/golden/encoder_remap.json by RUNNING THE REFERENCE's pretrained-weight loaders (build container only).

network/MSUNet.py:MSUNet.load_segface_weight (:63-160) and load_IMAGENET1K_weight (:162-240) of /root/reference are fed synthetic
checkpoints (oracle.msunet_oracle.encoder_checkpoint: tensor n is filled with n + 1, every model entry starts at -1); the fixture
records, per MS-UNet state_dict key, which checkpoint key the reference copied into it.  Depth 18 in stage 2 covers the
two-digit block indices.  The fixture travels; the reference does not.

    python oracle/make_remap_golden.py
"""
import json
import logging
import os
import sys
import tempfile
from types import SimpleNamespace as NS

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_shims"))
sys.path.insert(0, "/root/reference")

from oracle import msunet_oracle as O  # noqa: E402
from network.MSUNet import MSUNet  # noqa: E402  (reference)

DEPTHS, HEADS, EMBED, IMG = [2, 2, 18, 2], [1, 2, 4, 8], 32, 64


def config(**pre):
    return NS(MODEL=NS(SWIN=NS(PATCH_SIZE=4, IN_CHANS=3, EMBED_DIM=EMBED, DEPTHS=DEPTHS, NUM_HEADS=HEADS, WINDOW_SIZE=7, MLP_RATIO=4.0,
                               QKV_BIAS=True, APE=False, PATCH_NORM=True),
                       DROP_RATE=0.0, DROP_PATH_RATE=0.1, ATTN_DROP_RATE=0.0, **pre), TRAIN=NS(USE_CHECKPOINT=False))


out = {"depths": DEPTHS, "num_heads": HEADS, "embed_dim": EMBED, "img_size": IMG}
with tempfile.TemporaryDirectory() as tmp:
    for what, root, wrap in (("segface", "backbone.0.", "state_dict_backbone"), ("imagenet", "features.", None)):
        path = os.path.join(tmp, what + ".pth")
        cfg = config(PRETRAIN_SEGFACE=path, PRETRAIN_IMAGENET1K=path)
        m = MSUNet(cfg, img_size=IMG, num_classes=1)
        with torch.no_grad():
            for v in m.ms_unet.state_dict().values():
                v.fill_(-1)
        ckpt, _ = O.encoder_checkpoint(m.ms_unet.state_dict(), root)
        ckpt[root.split(".")[0] + ".1.classifier.weight" if what == "segface" else "head.weight"] = torch.zeros(3)
        by_id = {int(v.flatten()[0]): k for k, v in ckpt.items() if k.startswith(root)}
        torch.save({wrap: ckpt} if wrap else ckpt, path)
        (m.load_segface_weight if what == "segface" else m.load_IMAGENET1K_weight)(cfg, logging)
        out[what] = {k: (by_id[int(v.flatten()[0])] if float(v.flatten()[0]) > 0 else None) for k, v in m.ms_unet.state_dict().items()}
        print(what, "copied", sum(v is not None for v in out[what].values()), "of", len(out[what]))
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "encoder_remap.json"), "w"), indent=0, sort_keys=True)
