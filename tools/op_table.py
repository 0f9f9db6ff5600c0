"""Per-op table of one training step: launches, time, achieved GB/s (algorithmic bytes) and TFLOP/s by (kind, shape).
Each op is bracketed by CUDA events on the launch stream (analysis aid: the events serialise nothing, but an eager
step has launch gaps, so compare the SUM with the graph-replayed step time of bench.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import T96, synth_batch  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200 import ops  # noqa: E402
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
ops.PROF = []
step()
torch.cuda.synchronize()
agg = {}
for tag, a, b, nb, fl in ops.PROF:
    d = agg.setdefault(tag, [0.0, 0, 0, 0])
    d[0] += a.elapsed_time(b)
    d[1] += 1
    d[2] += nb
    d[3] += fl
ops.PROF = None
tot = sum(v[0] for v in agg.values())
print(f"sum of op times {tot:.2f} ms over {sum(v[1] for v in agg.values())} ops")
bykind = {}
for (M, N, K, kind), v in agg.items():
    k = bykind.setdefault(kind, [0.0, 0])
    k[0] += v[0]
    k[1] += v[1]
for kind, v in sorted(bykind.items(), key=lambda kv: -kv[1][0]):
    print(f"  {kind:16s} {v[0]:8.3f} ms  x{v[1]}")
print(f"{'kind':16s} {'M':>9s} {'N':>6s} {'K':>8s} {'n':>4s} {'ms':>8s} {'us/op':>8s} {'GB/s':>7s} {'TF/s':>7s}")
for (M, N, K, kind), v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    ms, n, nb, fl = v
    print(f"{kind:16s} {M:9d} {N:6d} {K:8d} {n:4d} {ms:8.3f} {1e3 * ms / n:8.1f} {nb / ms / 1e6:7.0f} {fl / ms / 1e9:7.1f}")
