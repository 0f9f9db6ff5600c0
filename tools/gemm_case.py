"""Run a few representative tcgen05 GEMM launches (for ncu / timing). Usage: gemm_case.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_segmentation_of_stylegan2_artifacts_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
bf = torch.bfloat16


def fc1(M, N, K):
    a = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) * 0.1).to(bf)
    b = torch.randn(N, device=dev)
    y = torch.empty(M, N, dtype=bf, device=dev)
    pre = torch.empty(M, N, dtype=bf, device=dev)
    return lambda: ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, Cpre=pre, bias=b, act=1), M, N, K, dev)


def plain(M, N, K):
    a = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) * 0.1).to(bf)
    y = torch.empty(M, N, dtype=bf, device=dev)
    return lambda: ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y), M, N, K, dev)


def conv(B, S, E):
    x = torch.randn(B * S * S, E, device=dev).to(bf)
    w = (torch.randn(E, 9 * E, device=dev) * 0.05).to(bf)
    b = torch.randn(E, device=dev)
    y = torch.empty(B * S * S, E, dtype=bf, device=dev)
    return lambda: ops.gemm(ops.operand(x, ld=E, map=ops.MAP_CONV3, geo=[S, S, E]), ops.operand(w), ops.epilogue(y, bias=b),
                            B * S * S, E, 9 * E, dev)


cases = {"fc1_s0": (fc1(262144, 384, 96), 2 * 262144 * 384 * 96, (262144 * 96 + 2 * 262144 * 384) * 2),
         "plain_s0": (plain(262144, 384, 96), 2 * 262144 * 384 * 96, (262144 * 96 + 262144 * 384) * 2),
         "fc1_s2": (fc1(16384, 1536, 384), 2 * 16384 * 1536 * 384, (16384 * 384 + 2 * 16384 * 1536) * 2),
         "plain_big": (plain(8192, 4096, 4096), 2 * 8192 * 4096 * 4096, 0),
         "conv_b4": (conv(4, 512, 96), 2 * 4 * 512 * 512 * 96 * 864, 2 * 4 * 512 * 512 * 96 * 2)}
for name, (fn, flops, byts) in cases.items():
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:10s} {ms * 1e3:9.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s  {byts / ms / 1e6:8.1f} GB/s")
