"""AdamW step over the T96 MS-UNet parameters (43 M fp32): fused multi-tensor kernel vs torch.optim.AdamW variants."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import T96  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.optim import FusedAdamW  # noqa: E402

dev = torch.device("cuda:0")
m = MSUNetSys(img_size=512, **T96).to(dev)
ps = [p for p in m.parameters()]
for p in ps:
    p.grad = torch.randn_like(p) * 0.01
n = sum(p.numel() for p in ps)
for name, mk in (("msu FusedAdamW", lambda: FusedAdamW(ps, lr=1e-4, weight_decay=0.01)),
                 ("torch AdamW fused=True", lambda: torch.optim.AdamW(ps, lr=1e-4, weight_decay=0.01, fused=True)),
                 ("torch AdamW foreach", lambda: torch.optim.AdamW(ps, lr=1e-4, weight_decay=0.01, foreach=True))):
    opt = mk()
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(10):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 10
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:24s}: device {ms * 1e3:8.1f} us/step  host wall {wall * 1e3:6.2f} ms/step  {28.0 * n / ms / 1e6:7.0f} GB/s of 28 B/param ({n / 1e6:.1f} M params, {len(ps)} tensors)")
