"""Time representative tcgen05 GEMM launches of the training step (CUDA-graph of `reps` launches, device timed).
Usage: gemm_case.py [reps] [case-name-substring ...]   (ncu: run with reps=1 and a case filter)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_segmentation_of_stylegan2_artifacts_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
filt = sys.argv[2:]
bf = torch.bfloat16


def fc1(M, N, K):
    a = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) * 0.1).to(bf)
    b = torch.randn(N, device=dev)
    y = torch.empty(M, N, dtype=bf, device=dev)
    pre = torch.empty(M, N, dtype=bf, device=dev)
    return (lambda: ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, Cpre=pre, bias=b, act=1), M, N, K, dev),
            2 * M * N * K, (M * K + 2 * M * N) * 2)


def plain(M, N, K):
    a = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) * 0.1).to(bf)
    y = torch.empty(M, N, dtype=bf, device=dev)
    return (lambda: ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y), M, N, K, dev),
            2 * M * N * K, (M * K + M * N) * 2)


def resid(M, N, K):
    a = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) * 0.1).to(bf)
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev).to(bf)
    y = torch.empty(M, N, dtype=bf, device=dev)
    return (lambda: ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, bias=b, R=r), M, N, K, dev),
            2 * M * N * K, (M * K + 2 * M * N) * 2)


def dgelu(M, N, K):
    a = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) * 0.1).to(bf)
    h = torch.randn(M, N, device=dev).to(bf)
    y = torch.empty(M, N, dtype=bf, device=dev)
    return (lambda: ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, H=h, ldh=N), M, N, K, dev),
            2 * M * N * K, (M * K + 2 * M * N) * 2)


def fc1g(M, N, K):
    """fc1 with the stored derivative (act = 2: C = GELU, Cpre = GELU')"""
    a = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) * 0.1).to(bf)
    b = torch.randn(N, device=dev)
    y = torch.empty(M, N, dtype=bf, device=dev)
    pre = torch.empty(M, N, dtype=bf, device=dev)
    return (lambda: ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, Cpre=pre, bias=b, act=2), M, N, K, dev),
            2 * M * N * K, (M * K + 2 * M * N) * 2)


def dmul(M, N, K):
    """dh with the stored derivative (act = 3: multiply by H)"""
    a = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) * 0.1).to(bf)
    h = torch.randn(M, N, device=dev).to(bf)
    y = torch.empty(M, N, dtype=bf, device=dev)
    return (lambda: ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, H=h, ldh=N, act=3), M, N, K, dev),
            2 * M * N * K, (M * K + 2 * M * N) * 2)


def conv(B, S, E):
    x = torch.randn(B * S * S, E, device=dev).to(bf)
    w = (torch.randn(E, 9 * E, device=dev) * 0.05).to(bf)
    b = torch.randn(E, device=dev)
    y = torch.empty(B * S * S, E, dtype=bf, device=dev)
    return (lambda: ops.gemm(ops.operand(x, ld=E, map=ops.MAP_CONV3, geo=[S, S, E]), ops.operand(w), ops.epilogue(y, bias=b),
                             B * S * S, E, 9 * E, dev), 2 * B * S * S * E * 9 * E, 2 * B * S * S * E * 2)


def conv_ep(B, S, E, kind):
    """the head's other conv launches: fwd = bias + GELU with a second output (act 1: pre-activation, act 2: GELU'),
    dgrad = x GELU'(H) (act 0: computed from the pre-activation, act 3: H holds the derivative), dgrad_un = the same through the
    inverse depth-to-space store"""
    x = torch.randn(B * S * S, E, device=dev).to(bf)
    w = (torch.randn(E, 9 * E, device=dev) * 0.05).to(bf)
    b = torch.randn(E, device=dev)
    y = torch.empty(B * S * S, E, dtype=bf, device=dev)
    aux = torch.randn(B * S * S, E, device=dev).to(bf)
    A = ops.operand(x, ld=E, map=ops.MAP_CONV3, geo=[S, S, E])
    M = B * S * S
    if kind.startswith("fwd"):
        ep = lambda: ops.epilogue(y, Cpre=aux, bias=b, act=int(kind[-1]))
    elif kind.startswith("dgrad_un"):
        r = S // 4
        y = torch.empty(B * r * r, 16 * E, dtype=bf, device=dev)
        ep = lambda: ops.epilogue(y, ldc=16 * E, H=aux, ldh=E, act=int(kind[-1]), map=ops.MAP_UNSHUFFLE, geo=[r, r, 4, E])
    else:
        ep = lambda: ops.epilogue(y, H=aux, ldh=E, act=int(kind[-1]))
    return (lambda: ops.gemm(A, ops.operand(w), ep(), M, E, 9 * E, dev), 2 * M * E * 9 * E, 3 * M * E * 2)


def conv_lnd(B, S, E):
    """second head conv with the LayerNorm + 1x1 conv fused into its epilogue (MsuEpilogue.lnd_*)"""
    x = torch.randn(B * S * S, E, device=dev).to(bf)
    w = (torch.randn(E, 9 * E, device=dev) * 0.05).to(bf)
    b = torch.randn(E, device=dev)
    y = torch.empty(B * S * S, E, dtype=bf, device=dev)
    g, be, ow = torch.ones(E, device=dev), torch.zeros(E, device=dev), torch.randn(E, device=dev)
    lo = torch.empty(B * S * S, dtype=bf, device=dev)
    st = torch.empty(3, B * S * S, device=dev)
    return (lambda: ops.gemm(ops.operand(x, ld=E, map=ops.MAP_CONV3, geo=[S, S, E]), ops.operand(w),
                             ops.epilogue(y, bias=b, lnd=(g, be, ow, lo, st[0], st[1], st[2])), B * S * S, E, 9 * E, dev),
            2 * B * S * S * E * 9 * E, 2 * B * S * S * E * 2)


def head_expand(Bn, r, E):
    M, N = Bn * r * r, 16 * E
    a = torch.randn(M, E, device=dev).to(bf)
    w = (torch.randn(N, E, device=dev) * 0.1).to(bf)
    y = torch.empty(16 * M, E, dtype=bf, device=dev)
    pre = torch.empty(16 * M, E, dtype=bf, device=dev)
    return (lambda: ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, ldc=E, Cpre=pre, act=1, map=ops.MAP_SHUFFLE, geo=[r, r, 4, E]),
                             M, N, E, dev), 2 * M * N * E, (M * E + 2 * M * N) * 2)


def proj_win(Bn, H, C):
    geo = [H, H, 7 * ((H + 6) // 7), 7 * ((H + 6) // 7), 3, 3]
    Tw, T = Bn * geo[2] * geo[3], Bn * H * H
    a = torch.randn(Tw, C, device=dev).to(bf)
    w = (torch.randn(C, C, device=dev) * 0.1).to(bf)
    b = torch.randn(C, device=dev)
    r = torch.randn(T, C, device=dev).to(bf)
    y = torch.empty(T, C, dtype=bf, device=dev)
    return (lambda: ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, bias=b, R=r, map=ops.MAP_WINDOW, geo=geo), Tw, C, C, dev),
            2 * Tw * C * C, (Tw * C + 2 * T * C) * 2)


cases = {
    "head_expand": lambda: head_expand(16, 128, 96), "projwin_s0": lambda: proj_win(16, 128, 96),
    "fc1_s0": lambda: fc1(262144, 384, 96), "fc2_s0": lambda: resid(262144, 96, 384),
    "dh_s0": lambda: dgelu(262144, 384, 96), "fc1g_s0": lambda: fc1g(262144, 384, 96), "dmul_s0": lambda: dmul(262144, 384, 96),
    "fc1g_s2": lambda: fc1g(16384, 1536, 384), "dmul_s2": lambda: dmul(16384, 1536, 384), "qkv_s0": lambda: plain(283024, 288, 96),
    "fc1_s1": lambda: fc1(65536, 768, 192), "fc2_s1": lambda: resid(65536, 192, 768),
    "fc1_s2": lambda: fc1(16384, 1536, 384), "plain_s2": lambda: plain(16384, 1536, 384),
    "dh_s2": lambda: dgelu(16384, 1536, 384), "fc2_s2": lambda: resid(16384, 384, 1536),
    "dxn_s2": lambda: plain(16384, 384, 1536), "qkv_s2": lambda: plain(19600, 1152, 384),
    "proj_s2": lambda: plain(19600, 384, 384), "fc1_s3": lambda: fc1(4096, 3072, 768),
    "convfwd1_b16": lambda: conv_ep(16, 512, 96, "fwd1"), "convfwd2_b16": lambda: conv_ep(16, 512, 96, "fwd2"),
    "convdgrad0_b16": lambda: conv_ep(16, 512, 96, "dgrad0"), "convdgrad3_b16": lambda: conv_ep(16, 512, 96, "dgrad3"),
    "convdgrad_un0_b16": lambda: conv_ep(16, 512, 96, "dgrad_un0"), "convdgrad_un3_b16": lambda: conv_ep(16, 512, 96, "dgrad_un3"),
    "plain_big": lambda: plain(8192, 4096, 4096), "conv_b16": lambda: conv(16, 512, 96), "convlnd_b16": lambda: conv_lnd(16, 512, 96),
}
for name, mk in cases.items():
    if filt and not any(f in name for f in filt):
        continue
    fn, flops, byts = mk()
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
    if reps > 1:
        g.replay()
    else:
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:10s} {ms * 1e3:9.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s  {byts / ms / 1e6:8.1f} GB/s", flush=True)
