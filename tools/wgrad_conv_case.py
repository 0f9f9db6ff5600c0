"""3x3 conv weight-gradient kernel timing at the head size (16 x 512 x 512 x 96)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_segmentation_of_stylegan2_artifacts_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, S, E = int(os.environ.get("B", 16)), 512, 96
Mp = B * S * S
x = torch.randn(Mp, E, device=dev).to(torch.bfloat16)
dz = torch.randn(Mp, E, device=dev).to(torch.bfloat16)
dw = torch.empty(E, 9 * E, device=dev)
db = torch.empty(E, device=dev)


def fn():
    ops.gemm(ops.operand(dz, orient=1), ops.operand(x, ld=E, orient=1, map=ops.MAP_CONV3, geo=[S, S, E]),
             ops.epilogue(dw, out_f32=True, colsum=None if os.environ.get('NOBIAS') else db), E, 9 * E, Mp, dev)


for _ in range(2):
    fn()
torch.cuda.synchronize()
ref = dw.clone()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"wgrad conv B={B}: {ms * 1e3:8.1f} us  {2.0 * Mp * E * 9 * E / ms / 1e9:7.1f} TFLOP/s  checksum {float(dw.double().abs().sum()):.6e}")
