"""nn.DataParallel (what the reference wraps the model in, trainer.py:96-97) over 2 GPUs vs the same step on one GPU."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import msunet_oracle as O  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys  # noqa: E402

cfg = O.Cfg(img_size=96, **O.T32)
sd = O.make_weights(cfg)
x, y = O.make_inputs(cfg, 4)
kw = dict(img_size=96, embed_dim=32, depths=[2, 2, 2, 2], num_heads=[1, 2, 4, 8], drop_path_rate=0.0)
res = {}
for name in ("single", "dataparallel"):
    m = MSUNetSys(**kw)
    m.load_state_dict(sd, strict=True)
    m = m.set_precision("fp32").to("cuda:0").train()
    net = torch.nn.DataParallel(m, device_ids=[0, 1]) if name == "dataparallel" else m
    logits = net(x.to("cuda:0"))
    loss = DynamicLoss(alpha=0.2, beta=0.8, tversky_bce_mix=0.45)(logits, y.to("cuda:0"))
    loss.backward()
    torch.cuda.synchronize()
    res[name] = (logits.detach().float().cpu(), loss.item(), m.layers[1].blocks[0].attn.qkv.weight.grad.cpu())
a, b = res["single"], res["dataparallel"]
e1 = float((a[0] - b[0]).abs().max() / a[0].abs().max())
e3 = float((a[2] - b[2]).abs().max() / a[2].abs().max())
print(f"nn.DataParallel vs single GPU: logits relmax {e1:.2e}  loss {a[1]:.6f} vs {b[1]:.6f}  grad relmax {e3:.2e}")
assert e1 < 1e-4 and abs(a[1] - b[1]) < 1e-5 and e3 < 1e-3
print("dp_smoke ok")
