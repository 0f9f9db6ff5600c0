"""One launch each of the LayerNorm kernels at the stage-0 size (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_segmentation_of_stylegan2_artifacts_b200 import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
T, C = 262144, 96
x = torch.randn(T, C, device=dev).to(bf); dy = torch.randn(T, C, device=dev).to(bf)
w = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
for _ in range(2):
    y, mean, rstd = ops.ln_fwd(x, w, b, T, C)
    ops.ln_bwd(dy, x, w, b, mean, rstd, T, C, dres=dy)
torch.cuda.synchronize()
