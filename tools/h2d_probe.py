"""Host -> device copy bandwidth of the box (pinned fp32, the e2e batch: 64 MiB), alone and while a training step runs."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import T96, synth_batch  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys  # noqa: E402

dev = torch.device("cuda:0")
x, y = synth_batch(16, 512, 1)
xh, yh = x.pin_memory(), y.pin_memory()
xd, yd = xh.to(dev), yh.to(dev)
cs = torch.cuda.Stream()


def copy_ms(n=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        e0.record(cs)
        for _ in range(n):
            xd.copy_(xh, non_blocking=True)
            yd.copy_(yh, non_blocking=True)
        e1.record(cs)
    return e0, e1, n


e0, e1, n = copy_ms()
torch.cuda.synchronize()
nbytes = (xh.numel() + yh.numel()) * 4
print(f"H2D alone: {e0.elapsed_time(e1) / n:.2f} ms per 64 MiB batch = {nbytes / (e0.elapsed_time(e1) / n) / 1e6:.1f} GB/s")
m = MSUNetSys(img_size=512, drop_path_rate=0.1, **T96).to(dev).train()
crit = DynamicLoss(alpha=0.2, beta=0.8, tversky_bce_mix=0.45)
xs, ys = xd.clone(), yd.clone()


def step():
    for p in m.parameters():
        p.grad = None
    crit(m(xs), ys).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
g.replay()
torch.cuda.synchronize()
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
for _ in range(5):
    g.replay()
s1.record()
torch.cuda.synchronize()
print(f"step alone: {s0.elapsed_time(s1) / 5:.2f} ms")
s0.record()
e0, e1, n = copy_ms(5)
for _ in range(5):
    g.replay()
s1.record()
torch.cuda.synchronize()
print(f"5 steps with 5 concurrent batch copies: step {s0.elapsed_time(s1) / 5:.2f} ms, copy {e0.elapsed_time(e1) / n:.2f} ms per batch "
      f"= {nbytes / (e0.elapsed_time(e1) / n) / 1e6:.1f} GB/s")
