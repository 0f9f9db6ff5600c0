"""Per-kernel time breakdown of one training step (torch.profiler/CUPTI; analysis aid, not a bench)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import T96, synth_batch  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys  # noqa: E402

B = int(os.environ.get("B", 16))
S = int(os.environ.get("S", 512))
dev = torch.device("cuda:0")
m = MSUNetSys(img_size=S, drop_path_rate=float(os.environ.get("DP", 0.1)), **T96).to(dev).train()
crit = DynamicLoss(alpha=0.2, beta=0.8, tversky_bce_mix=0.45)
x, y = synth_batch(B, S, 1)
x, y = x.to(dev), y.to(dev)


def step():
    for p in m.parameters():
        p.grad = None
    crit(m(x), y).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name.split("(")[0][:110]
        r = rows.setdefault(name, [0.0, 0])
        r[0] += e.device_time / 1e3 if hasattr(e, "device_time") else e.cuda_time / 1e3
        r[1] += 1
tot = sum(v[0] for v in rows.values())
print(f"total kernel time {tot:.2f} ms over {sum(v[1] for v in rows.values())} launches")
for k, v in sorted(rows.items(), key=lambda kv: -kv[1][0])[:32]:
    print(f"{v[0]:9.3f} ms {100 * v[0] / tot:5.1f}%  x{v[1]:4d}  {k}")
