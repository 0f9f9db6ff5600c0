"""Time the window-attention core (forward / backward kernels) at the four stage shapes of T96 @ 512^2, B=16.
Usage: attn_case.py [reps] [stage ...]   (MSU_ATT_TRACE=1 with reps=1 dumps the backward phase timeline)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_segmentation_of_stylegan2_artifacts_b200 import ops  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.functional import window_geo  # noqa: E402

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
filt = sys.argv[2:]
bf = torch.bfloat16
Bn = int(os.environ.get("B", 16))
S = int(os.environ.get("S", 512))
for k in range(4):
    if filt and str(k) not in filt:
        continue
    H = S // (4 << k)
    C = 96 << k
    nH = C // 32
    geo = window_geo(H, H, 3)
    nW = Bn * (geo[2] // 7) * (geo[3] // 7)
    qkv = (torch.randn(nW * 49, 3 * C, device=dev) * 0.5).to(bf)
    bias = ops.relbias_expand(torch.randn(169, nH, device=dev) * 0.1, nH)
    do = torch.randn(nW * 49, C, device=dev).to(bf)
    o, lse = ops.winattn_fwd(qkv, bias, nW, nH, geo, want_lse=True)
    use_lse = os.environ.get("LSE", "1") != "0"      # the training path hands the forward's log-sum-exp to the backward
    fns = {"fwd": (lambda: ops.winattn_fwd(qkv, bias, nW, nH, geo, want_lse=True), (qkv.numel() + o.numel()) * 2),
           "bwd": (lambda: ops.winattn_bwd(qkv, bias, o, do, nW, nH, geo, lse=lse if use_lse else None), (2 * qkv.numel() + do.numel()) * 2)}
    for name, (fn, byts) in fns.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if reps > 1:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(reps):
                    fn()
            g.replay()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay() if reps > 1 else fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"stage {k} {name}: windows {nW:6d} heads {nH:2d}  {ms * 1e3:8.1f} us  {byts / ms / 1e6:7.0f} GB/s (algorithmic)", flush=True)
