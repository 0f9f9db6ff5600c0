"""LayerNorm / colsum / gather micro-timings (GB/s against the HBM roofline)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_segmentation_of_stylegan2_artifacts_b200 import ops  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.functional import window_geo  # noqa: E402

dev = torch.device("cuda:0")
bf = torch.bfloat16


def timeit(fn, reps=10):
    """Device time per call: `reps` calls captured in one CUDA graph (no CPU launch overhead in the number)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (B, H, C) in [(16, 128, 96), (16, 64, 192), (16, 32, 384), (16, 16, 768)]:
    T = B * H * H
    x = torch.randn(T, C, device=dev).to(bf)
    dy = torch.randn(T, C, device=dev).to(bf)
    w = torch.ones(C, device=dev)
    b = torch.zeros(C, device=dev)
    y, mean, rstd = ops.ln_fwd(x, w, b, T, C)
    ms = timeit(lambda: ops.ln_fwd(x, w, b, T, C))
    print(f"ln_fwd plain   T={T} C={C}: {ms*1e3:7.1f} us  {2*T*C*2/ms/1e6:7.0f} GB/s")
    ms = timeit(lambda: ops.ln_bwd(dy, x, w, b, mean, rstd, T, C, dres=dy))
    print(f"ln_bwd plain+r T={T} C={C}: {ms*1e3:7.1f} us  {4*T*C*2/ms/1e6:7.0f} GB/s (incl. param reduce)")
    geo = window_geo(H, H, 3)
    nW = (geo[2] // 7) * (geo[3] // 7)
    Tw = B * nW * 49
    ms = timeit(lambda: ops.ln_fwd(x, w, b, Tw, C, out_map=ops.MAP_WINDOW, geo=geo, n_stat_rows=T))
    print(f"ln_fwd window  T={T} C={C}: {ms*1e3:7.1f} us  {(T+Tw)*C*2/ms/1e6:7.0f} GB/s")
    dyw = torch.randn(Tw, C, device=dev).to(bf)
    ms = timeit(lambda: ops.ln_bwd(dyw, x, w, b, mean, rstd, T, C, dres=dy, dy_map=ops.MAP_WINDOW, geo=geo))
    print(f"ln_bwd window  T={T} C={C}: {ms*1e3:7.1f} us  {4*T*C*2/ms/1e6:7.0f} GB/s")
    ms = timeit(lambda: ops.colsum(ops.operand(x), T, C, dev))
    print(f"colsum         T={T} C={C}: {ms*1e3:7.1f} us  {T*C*2/ms/1e6:7.0f} GB/s")
    ms = timeit(lambda: ops.gather_rows(ops.operand(x, map=ops.MAP_WINDOW, geo=geo), Tw, C, x))
    print(f"gather window  T={T} C={C}: {ms*1e3:7.1f} us  {(T+Tw)*C*2/ms/1e6:7.0f} GB/s")

# head: LayerNorm(96) fused with the 1x1 conv, 16 x 512 x 512 pixels
T, C = 16 * 512 * 512, 96
x = torch.randn(T, C, device=dev).to(bf)
w = torch.ones(C, device=dev)
b = torch.zeros(C, device=dev)
ow = torch.randn(C, device=dev)
y, mean, rstd = ops.ln_fwd(x, w, b, T, C, dotw=ow)
ms = timeit(lambda: ops.ln_fwd(x, w, b, T, C, dotw=ow), 3)
print(f"ln_fwd head    T={T} C={C}: {ms*1e3:7.1f} us  {(T*C*2+T*2)/ms/1e6:7.0f} GB/s")
dl = torch.randn(T, device=dev).to(bf)
ms = timeit(lambda: ops.ln_bwd(dl, x, w, b, mean, rstd, T, C, dotw=ow), 3)
print(f"ln_bwd head    T={T} C={C}: {ms*1e3:7.1f} us  {(2*T*C*2+T*2)/ms/1e6:7.0f} GB/s")
