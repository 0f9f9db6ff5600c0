"""Weight-gradient GEMM timings (tcgen05 MN-major split-K kernel + reduce)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_segmentation_of_stylegan2_artifacts_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf = torch.bfloat16
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
only = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]]
for (T, I, J) in only or [(262144, 384, 96), (283024, 288, 96), (262144, 96, 384), (65536, 768, 192), (78400, 576, 192),
                  (16384, 1536, 384), (19600, 1152, 384), (16384, 384, 1536), (4096, 3072, 768), (7056, 2304, 768)]:
    dy = torch.randn(T, I, device=dev).to(bf)
    x = torch.randn(T, J, device=dev).to(bf)
    dw = torch.empty(I, J, device=dev)

    def fn():
        ops.gemm(ops.operand(dy, orient=1), ops.operand(x, orient=1), ops.epilogue(dw, out_f32=True), I, J, T, dev)
    for _ in range(2):
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
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * T * I * J
    by = T * (I + J) * 2
    print(f"wgrad T={T:7d} I={I:5d} J={J:5d}: {ms*1e3:7.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  {by/ms/1e6:7.0f} GB/s")
