"""Thin, allocation-explicit wrappers over the C ABI (one Python function per kernel entry).

Every function enqueues on torch's current CUDA stream and returns immediately.  Tensors are
plain torch tensors used as device buffers; no torch math happens here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L
from ._lib import (BF16, F16, F32, MAP_CONV3, MAP_MERGE, MAP_NONE, MAP_SHUFFLE, MAP_UNSHUFFLE, MAP_WINDOW,  # noqa: F401
                   MsuEpilogue, MsuOperand)

_ws_cache: dict = {}
GEMM_BACKEND = 0  # 0 auto (tcgen05 where supported), 1 force SIMT


def _need_cuda(t: torch.Tensor, name: str = "tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device: the MS-UNet hot path has no CPU fallback")


def workspace(device: torch.device, elems: int = 16 << 20) -> torch.Tensor:
    """Per-device fp32 scratch for deterministic split-K / column-sum partials."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < elems:
        ws = torch.empty(elems, dtype=torch.float32, device=device)
        _ws_cache[key] = ws
    return ws


def operand(t: torch.Tensor, ld: Optional[int] = None, *, t2: Optional[torch.Tensor] = None, ld2: int = 0,
            k_split: int = 0, orient: int = 0, map: int = MAP_NONE, geo: Optional[Sequence[int]] = None,
            rowscale: Optional[torch.Tensor] = None, rps: int = 0, offset: int = 0) -> MsuOperand:
    o = MsuOperand()
    o.ptr = t.data_ptr() + offset * t.element_size()
    o.ptr2 = L.ptr(t2)
    o.ld = t.shape[-1] if ld is None else ld
    o.ld2 = ld2
    o.k_split = k_split
    o.orient = orient
    o.map = map
    o.dtype = L.dt(t)
    o.geo = L.geo6(geo)
    o.rowscale = L.ptr(rowscale)
    o.rows_per_sample = rps
    o._keep = (t, t2, rowscale)
    return o


def epilogue(Cm: torch.Tensor, ldc: Optional[int] = None, *, Cpre=None, bias=None, R=None, ldr: Optional[int] = None,
             H=None, ldh: int = 0, rowscale=None, rps: int = 0, act: int = 0, map: int = MAP_NONE, geo=None,
             out_f32: bool = False, accumulate: bool = False, offset: int = 0, colsum=None, lnd=None,
             dtype: Optional[torch.dtype] = None) -> MsuEpilogue:
    """`lnd` = (gamma, beta, w, logits, mean, rstd, m2): fused head LayerNorm + 1x1 conv (see MsuEpilogue.lnd_*); `Cm` may then be
    None (the rows are not stored; `ldc` and `dtype` must be given)."""
    e = MsuEpilogue()
    e.C = None if Cm is None else Cm.data_ptr() + offset * Cm.element_size()
    if Cm is None:
        if lnd is None or ldc is None or dtype is None:
            raise ValueError("an output tensor is required (only the fused LayerNorm + dot epilogue may drop it)")
        e.Cpre = e.bias = e.R = e.H = None
        e.bias = L.ptr(bias)
        e.ldc = e.ldr = ldc
        e.ldh = 0
        e.rowscale = None
        e.rows_per_sample = 0
        e.act = 0
        e.map = MAP_NONE
        e.dtype = L._DT[dtype]
        e.geo = L.geo6(None)
        e.out_f32 = e.accumulate = 0
        e.colsum = None
        (e.lnd_gamma, e.lnd_beta, e.lnd_w, e.lnd_logits, e.lnd_mean, e.lnd_rstd, e.lnd_m2) = (L.ptr(v) for v in lnd)
        return e
    e.Cpre = L.ptr(Cpre)
    e.bias = L.ptr(bias)
    e.R = L.ptr(R)
    e.H = L.ptr(H)
    e.ldc = Cm.shape[-1] if ldc is None else ldc
    e.ldr = e.ldc if ldr is None else ldr
    e.ldh = ldh
    e.rowscale = L.ptr(rowscale)
    e.rows_per_sample = rps
    e.act = act
    e.map = map
    e.dtype = F32 if out_f32 else L.dt(Cm)
    e.geo = L.geo6(geo)
    e.out_f32 = 1 if out_f32 else 0
    e.accumulate = 1 if accumulate else 0
    e.colsum = L.ptr(colsum)
    if lnd is not None:
        (e.lnd_gamma, e.lnd_beta, e.lnd_w, e.lnd_logits, e.lnd_mean, e.lnd_rstd, e.lnd_m2) = (L.ptr(v) for v in lnd)
    if colsum is not None and (colsum.dtype != torch.float32 or not out_f32):
        raise TypeError("colsum (bias gradient) needs an fp32 vector and an fp32 weight-gradient output")
    if out_f32 and Cm.dtype != torch.float32:
        raise TypeError("out_f32 needs an fp32 output tensor")
    for t in (Cpre, R, H):
        if t is not None and t.dtype != Cm.dtype:
            raise TypeError("epilogue tensors must share the output dtype")
    e._keep = (Cm, Cpre, bias, R, H, rowscale, colsum)
    return e


import threading as _threading

_tls = _threading.local()   # per-thread: nn.DataParallel runs one autograd thread per device


def set_side(fork):
    """functional.py: the active weight-gradient fork of this thread (parameter-gradient reduces ride on its side stream)."""
    prev = getattr(_tls, "side", None)
    _tls.side = fork
    return prev


def _off_path(fn, *keep):
    """Run a launch whose result is only needed after the backward joins (parameter-gradient reductions).
    `keep`: temporaries read by that launch; they must outlive the join (the caching allocator would otherwise hand
    their memory to a later main-stream allocation while the side stream still reads it)."""
    side = getattr(_tls, "side", None)
    if side is not None:
        side.keep.extend(keep)
        side.run(fn)
    else:
        fn()


# ----------------------------------------------------------------------------------------------
# Gradient slots: where a parameter's gradient should be WRITTEN by the kernel that produces it.  dp.DataParallelB200 registers
# one fp32 view of its flat all-reduce bucket per parameter (keyed by the parameter's data pointer); the backward Functions ask
# `grad_out(param, shape)` instead of torch.empty, so the weight-gradient GEMM / reduce writes straight into the bucket and the
# per-parameter copy into it disappears.  A slot is handed out at most once per backward pass and only while the parameter has no
# accumulated .grad (shared weights and gradient accumulation get private tensors and autograd adds them up as usual).
# ----------------------------------------------------------------------------------------------
_grad_slots: dict = {}      # param.data_ptr() -> [view, weakref(param), generation last handed out]
_grad_gen = [0]


def register_grad_slots(slots: dict) -> None:
    """slots: {parameter: fp32 tensor view with the parameter's numel}.  Replaces the registrations of those parameters."""
    import weakref
    mine = {id(prm) for prm in slots}
    for key in [k for k, e in _grad_slots.items() if e[1]() is None or id(e[1]()) in mine]:
        del _grad_slots[key]                    # parameters may have moved (data re-pointed at a flat buffer): re-key them
    for prm, view in slots.items():
        _grad_slots[prm.data_ptr()] = [view, weakref.ref(prm), -1]


def clear_grad_slots(params=None) -> None:
    if params is None:
        _grad_slots.clear()
    else:
        for prm in params:
            _grad_slots.pop(prm.data_ptr(), None)


def next_grad_pass() -> None:
    """Called once per optimisation step (after the gradients were consumed): slots may be handed out again."""
    _grad_gen[0] += 1


def grad_out(param: torch.Tensor, shape=None) -> torch.Tensor:
    """fp32 output tensor for the gradient of `param` (shape defaults to the parameter's)."""
    shape = tuple(param.shape) if shape is None else tuple(shape)
    ent = _grad_slots.get(param.data_ptr())
    if ent is not None and ent[2] != _grad_gen[0]:
        prm = ent[1]()
        if prm is not None and prm.grad is None and ent[0].numel() == param.numel() and ent[0].device == param.device:
            ent[2] = _grad_gen[0]
            return ent[0].view(shape)
    return torch.empty(shape, dtype=torch.float32, device=param.device)


PROF = None  # bench.py / tools set this to a list to collect (tag, start_event, end_event, bytes, flops) per launch


def _p0():
    if PROF is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _p1(e0, tag, nbytes=0, flops=0):
    """tag = (M, N, K, kind); nbytes = algorithmic HBM bytes of the launch."""
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    PROF.append((tag, e0, e1, nbytes, flops))


def gemm(A: MsuOperand, B: MsuOperand, E: MsuEpilogue, M: int, N: int, K: int, dev: torch.device) -> None:
    ws = workspace(dev)
    e0 = _p0()
    L.check(L.lib().msu_gemm(C.byref(A), C.byref(B), C.byref(E), M, N, K, ws.data_ptr(), ws.numel(), GEMM_BACKEND,
                             L.stream_ptr()), "msu_gemm")
    if e0 is not None:
        es = 2 if A.dtype == BF16 else 4
        if A.orient == 1:      # weight gradient: both [tokens, channels] operands read once, fp32 output
            kind = "wgrad"
            nb = K * (M + (N // 9 if B.map == MAP_CONV3 else N)) * es + M * N * 4
        else:
            kind = "conv3x3" if A.map == MAP_CONV3 else "gemm"
            if E.lnd_w is not None:
                kind += "_lnd"             # its own template instantiation: LayerNorm + dot statistics in the epilogue
            a_cols = K // 9 if A.map == MAP_CONV3 else K
            outs = 1 + (E.Cpre is not None) + (E.R is not None) + (E.H is not None)
            nb = (M * a_cols + N * K + outs * M * N) * es
        kind += "_tc" if L.lib().msu_last_gemm_backend() == 1 else "_simt"
        _p1(e0, (M, N, K, kind), nb, 2 * M * N * K)


def colsum(X: MsuOperand, M: int, N: int, dev: torch.device) -> torch.Tensor:
    out = torch.empty(N, dtype=torch.float32, device=dev)
    ws = workspace(dev)
    e0 = _p0()
    L.check(L.lib().msu_colsum(C.byref(X), M, N, out.data_ptr(), 0, ws.data_ptr(), ws.numel(), L.stream_ptr()),
            "msu_colsum")
    _p1(e0, (M, N, 0, "colsum"), M * N * (2 if X.dtype == BF16 else 4))
    return out


def _geo_arr(geo):
    return None if geo is None else L.geo6(geo)


def ln_fwd(x: torch.Tensor, gamma, beta, rows: int, Cdim: int, *, in_map=MAP_NONE, out_map=MAP_NONE, geo=None,
           n_stat_rows: Optional[int] = None, dotw: Optional[torch.Tensor] = None):
    """Returns (y, mean, rstd).  y is [rows, C] (or [rows] logits when dotw is given; rstd is then [2, rows]: the second row
    holds the per-row dot statistic the backward needs, see msu_ln_fwd)."""
    _need_cuda(x, "x")
    y = torch.empty((rows,) if dotw is not None else (rows, Cdim), dtype=x.dtype, device=x.device)
    ns = rows if n_stat_rows is None else n_stat_rows
    mean = torch.empty(ns, dtype=torch.float32, device=x.device)
    rstd = torch.empty((2, ns) if dotw is not None else (ns,), dtype=torch.float32, device=x.device)
    g = _geo_arr(geo)
    e0 = _p0()
    L.check(L.lib().msu_ln_fwd(L.dt(x), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
                               mean.data_ptr(), rstd.data_ptr(), rows, Cdim, in_map, out_map,
                               None if g is None else C.cast(g, C.c_void_p), L.ptr(dotw),
                               rstd[1].data_ptr() if dotw is not None else None, L.stream_ptr()), "msu_ln_fwd")
    _p1(e0, (rows, Cdim, 0, "ln_fwd" + ("_dot" if dotw is not None else "")),
        (ns * Cdim + y.numel()) * x.element_size())
    return y, mean, rstd


def ln_bwd(dy: torch.Tensor, x: torch.Tensor, gamma, beta, mean, rstd, rows: int, Cdim: int, *, dres=None,
           dy_map=MAP_NONE, dx_map=MAP_NONE, geo=None, dotw=None, dx_shape=None):
    """Returns (dx, dgamma, dbeta, ddotw|None); rows = number of LayerNorm rows."""
    dev = x.device
    dx = torch.empty(x.shape if dx_shape is None else dx_shape, dtype=x.dtype, device=dev)
    P = L.lib().msu_ln_bwd_partial_rows(L.dt(x), rows, Cdim)
    part = torch.empty(P * 3 * Cdim, dtype=torch.float32, device=dev)
    g = _geo_arr(geo)
    e0 = _p0()
    L.check(L.lib().msu_ln_bwd(L.dt(x), dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                               mean.data_ptr(), rstd.data_ptr(), L.ptr(dres), dx.data_ptr(), rows, Cdim, dy_map,
                               dx_map, None if g is None else C.cast(g, C.c_void_p), L.ptr(dotw),
                               rstd[1].data_ptr() if dotw is not None else None, part.data_ptr(),
                               L.stream_ptr()), "msu_ln_bwd")
    dg, db = grad_out(gamma), grad_out(beta)
    dw = grad_out(dotw) if dotw is not None else None
    _off_path(lambda: L.check(L.lib().msu_ln_param_reduce(part.data_ptr(), P, Cdim, dg.data_ptr(), db.data_ptr(), L.ptr(dw), 0,
                                                          L.stream_ptr()), "msu_ln_param_reduce"), part)
    _p1(e0, (rows, Cdim, 0, "ln_bwd" + ("_dot" if dotw is not None else "") + ("_res" if dres is not None else "")),
        (dy.numel() + (2 + (dres is not None)) * rows * Cdim) * x.element_size())
    return dx, dg, db, dw


_win_bufs: dict = {}


def window_rows_buffer(Tw: int, Cdim: int, geo, like: torch.Tensor) -> torch.Tensor:
    """Persistent [Tw, C] buffer for window-ordered gradient rows of one (geometry, width): its padding rows are zeroed once
    and never written again (msu_ln_bwd_dual only writes real pixels), so no per-call memset or gather is needed.  Reuse is
    stream-safe because every SwinBlockFn.backward joins its side stream before the next block's backward starts."""
    key = (like.device.index, like.dtype, Tw, Cdim, tuple(geo), torch.cuda.current_stream(like.device).cuda_stream)
    buf = _win_bufs.get(key)
    if buf is None:
        buf = torch.zeros(Tw, Cdim, dtype=like.dtype, device=like.device)
        _win_bufs[key] = buf
    return buf


def ln_bwd_dual(dy: torch.Tensor, x: torch.Tensor, gamma, beta, mean, rstd, rows: int, Cdim: int, dres, dxw: torch.Tensor,
                wgeo, rowscale, rps: int):
    """LayerNorm backward that also writes rowscale * dx into the window-ordered rows of `dxw`.  Returns (dx, dgamma, dbeta)."""
    dev = x.device
    dx = torch.empty_like(x)
    P = L.lib().msu_ln_bwd_partial_rows(L.dt(x), rows, Cdim)
    part = torch.empty(P * 3 * Cdim, dtype=torch.float32, device=dev)
    g = L.geo6(wgeo)
    e0 = _p0()
    L.check(L.lib().msu_ln_bwd_dual(L.dt(x), dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(),
                                    rstd.data_ptr(), L.ptr(dres), dx.data_ptr(), dxw.data_ptr(), rows, Cdim,
                                    C.cast(g, C.c_void_p), L.ptr(rowscale), int(rps), part.data_ptr(), L.stream_ptr()),
            "msu_ln_bwd_dual")
    dg, db = grad_out(gamma), grad_out(beta)
    _off_path(lambda: L.check(L.lib().msu_ln_param_reduce(part.data_ptr(), P, Cdim, dg.data_ptr(), db.data_ptr(), None, 0,
                                                          L.stream_ptr()), "msu_ln_param_reduce"), part)
    _p1(e0, (rows, Cdim, 0, "ln_bwd_dual"), (dy.numel() + 4 * rows * Cdim) * x.element_size())
    return dx, dg, db


def relbias_expand(table: torch.Tensor, nH: int) -> torch.Tensor:
    bias = torch.empty(nH, 49, 49, dtype=torch.float32, device=table.device)
    L.check(L.lib().msu_relbias_expand(table.data_ptr(), bias.data_ptr(), nH, L.stream_ptr()), "msu_relbias_expand")
    return bias


def winattn_fwd(qkv: torch.Tensor, bias: torch.Tensor, n_windows: int, nH: int, geo, p_drop: float = 0.0,
                seed: Optional[torch.Tensor] = None, want_lse: bool = False):
    """`seed`: int32[2] device tensor (two 32-bit words of the dropout counter hash); the backward needs the same.
    `want_lse`: also return the rows' log2-domain log-sum-exp [n_windows * 49, nH] fp32 for `winattn_bwd(lse=...)`."""
    o = torch.empty(n_windows * 49, nH * 32, dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty(n_windows * 49, nH, dtype=torch.float32, device=qkv.device) if want_lse else None
    g = L.geo6(geo)
    e0 = _p0()
    L.check(L.lib().msu_winattn_fwd(L.dt(qkv), qkv.data_ptr(), bias.data_ptr(), o.data_ptr(), n_windows, nH,
                                    C.cast(g, C.c_void_p), float(p_drop), L.ptr(seed), L.ptr(lse), L.stream_ptr()), "msu_winattn_fwd")
    _p1(e0, (n_windows, nH, 0, "winattn_fwd"), (qkv.numel() + o.numel()) * qkv.element_size(),
        2 * 2 * n_windows * nH * 49 * 49 * 32)
    return (o, lse) if want_lse else o


def winattn_bwd(qkv, bias, o, do, n_windows: int, nH: int, geo, p_drop: float = 0.0, seed=None, table=None, lse=None):
    """Returns (dqkv, dtable[169, nH]).  `table`: the bias-table parameter (its gradient slot is used when one is registered);
    `lse`: the forward's log-sum-exp (`winattn_fwd(want_lse=True)`), or None (the kernel recomputes the row statistics)."""
    dev = qkv.device
    dqkv = torch.empty_like(qkv)
    # direct: the kernel adds the bias-table gradient into dtable itself (atomics); else per-CTA partial slabs + a fixed-order reduce
    direct = L.lib().msu_winattn_bwd_direct(L.dt(qkv)) == 1
    dtable = grad_out(table, (169, nH)) if table is not None else torch.empty(169, nH, dtype=torch.float32, device=dev)
    part = None
    if direct:
        dtable.zero_()
    else:
        gx = L.lib().msu_winattn_bwd_grid(L.dt(qkv), n_windows, nH)
        part = torch.empty(gx * nH * 2401, dtype=torch.float32, device=dev)
    g = L.geo6(geo)
    e0 = _p0()
    L.check(L.lib().msu_winattn_bwd(L.dt(qkv), qkv.data_ptr(), bias.data_ptr(), o.data_ptr(), do.data_ptr(),
                                    dqkv.data_ptr(), L.ptr(part), dtable.data_ptr() if direct else None, n_windows, nH,
                                    C.cast(g, C.c_void_p), float(p_drop), L.ptr(seed), L.ptr(lse), L.stream_ptr()), "msu_winattn_bwd")
    if not direct:
        _off_path(lambda: L.check(L.lib().msu_relbias_reduce(part.data_ptr(), gx, nH, dtable.data_ptr(), 0, L.stream_ptr()),
                                  "msu_relbias_reduce"), part)
    _p1(e0, (n_windows, nH, 0, "winattn_bwd"), (2 * qkv.numel() + 2 * o.numel()) * qkv.element_size(),
        5 * 2 * n_windows * nH * 49 * 49 * 32)
    return dqkv, dtable


def prep_weight(mode: int, src: torch.Tensor, R: int, Cc: int, out_shape, dtype: torch.dtype) -> torch.Tensor:
    out = torch.empty(out_shape, dtype=dtype, device=src.device)
    L.check(L.lib().msu_prep_weight(mode, L.dt(out), src.data_ptr(), out.data_ptr(), R, Cc, L.stream_ptr()),
            "msu_prep_weight")
    return out


def prep_weight_into(mode: int, src: torch.Tensor, out: torch.Tensor, R: int, Cc: int) -> None:
    L.check(L.lib().msu_prep_weight(mode, L.dt(out), src.data_ptr(), out.data_ptr(), R, Cc, L.stream_ptr()),
            "msu_prep_weight")


def patchify4(img: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    B, _, S, _ = img.shape
    out = torch.empty(B * (S // 4) ** 2, 64, dtype=dtype, device=img.device)
    L.check(L.lib().msu_patchify4(L.dt(out), img.data_ptr(), out.data_ptr(), B, S, L.stream_ptr()), "msu_patchify4")
    return out


def stage_u8(img_hwc: torch.Tensor, label: torch.Tensor | None = None, flip: torch.Tensor | None = None):
    """dataset/dataset.py:13-16, 49-63 for a batch on the device: uint8 [B,H,W,3] (+ uint8 [B,H,W] label, + uint8 [B] flip flags) ->
    float32 [B,3,H,W] image / 255 and float32 [B,H,W] (label > 127)."""
    if img_hwc.dtype != torch.uint8 or img_hwc.dim() != 4 or img_hwc.shape[-1] != 3:
        raise ValueError(f"stage_u8: image must be uint8 [B,H,W,3], got {img_hwc.dtype} {tuple(img_hwc.shape)}")
    _need_cuda(img_hwc, "stage_u8 image")
    B, H, W, _ = img_hwc.shape
    img_hwc = img_hwc.contiguous()
    if label is not None:
        if label.dtype != torch.uint8 or tuple(label.shape) != (B, H, W):
            raise ValueError(f"stage_u8: label must be uint8 [{B},{H},{W}], got {label.dtype} {tuple(label.shape)}")
        label = label.contiguous()
    if flip is not None:
        if flip.dtype not in (torch.uint8, torch.bool) or flip.numel() != B:
            raise ValueError("stage_u8: flip must be uint8 / bool [B]")
        flip = flip.contiguous().view(torch.uint8)
    out = torch.empty(B, 3, H, W, dtype=torch.float32, device=img_hwc.device)
    lab = torch.empty(B, H, W, dtype=torch.float32, device=img_hwc.device) if label is not None else None
    if B == 0:
        return out, lab
    L.check(L.lib().msu_stage_u8(img_hwc.data_ptr(), label.data_ptr() if label is not None else None,
                                 flip.data_ptr() if flip is not None else None, out.data_ptr(),
                                 lab.data_ptr() if lab is not None else None, B, H, W, L.stream_ptr()), "msu_stage_u8")
    return out, lab


def gather_rows(src: MsuOperand, M: int, N: int, like: torch.Tensor) -> torch.Tensor:
    """Dense [M, N] copy of a mapped / scaled operand (same dtype as `like`)."""
    out = torch.empty(M, N, dtype=like.dtype, device=like.device)
    e0 = _p0()
    L.check(L.lib().msu_gather_rows(C.byref(src), out.data_ptr(), M, N, L.stream_ptr()), "msu_gather_rows")
    _p1(e0, (M, N, 0, "gather_rows"), 2 * M * N * like.element_size())
    return out


def cast(src: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if src.dtype == dtype:
        return src
    src = src.contiguous()
    out = torch.empty(src.shape, dtype=dtype, device=src.device)
    L.check(L.lib().msu_cast(L.dt(src), L.dt(out), src.data_ptr(), out.data_ptr(), src.numel(), L.stream_ptr()),
            "msu_cast")
    return out


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(a)
    L.check(L.lib().msu_add(L.dt(a), a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), L.stream_ptr()), "msu_add")
    return y


def loss_fwd(logits: torch.Tensor, target: torch.Tensor, alpha: float, beta: float, mix: float):
    """logits [B, N] (any of f32/bf16/f16), target [B, N] fp32 -> (loss[1], stats[B,8], flag[1])."""
    B, N = logits.shape
    dev = logits.device
    ws = torch.empty(B * 64 * 16, dtype=torch.float32, device=dev)
    stats = torch.empty(B, 8, dtype=torch.float32, device=dev)
    flag = torch.empty(1, dtype=torch.int32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    L.check(L.lib().msu_loss_fwd(L.dt(logits), logits.data_ptr(), target.data_ptr(), B, N, alpha, beta, mix,
                                 ws.data_ptr(), stats.data_ptr(), flag.data_ptr(), loss.data_ptr(), L.stream_ptr()),
            "msu_loss_fwd")
    return loss, stats, flag


def loss_per_sample(logits: torch.Tensor, target: torch.Tensor, alpha: float, beta: float, mix: float) -> torch.Tensor:
    """Per-image losses [B]: every image evaluated as its own batch of one, with its own {0,255} label decision."""
    B, N = logits.shape
    dev = logits.device
    ws = torch.empty(B * 64 * 16, dtype=torch.float32, device=dev)
    stats = torch.empty(B, 8, dtype=torch.float32, device=dev)
    L.check(L.lib().msu_loss_per_sample(L.dt(logits), logits.data_ptr(), target.data_ptr(), B, N, alpha, beta, mix,
                                        ws.data_ptr(), stats.data_ptr(), L.stream_ptr()), "msu_loss_per_sample")
    return stats[:, 4].clone()


def loss_bwd(logits, target, alpha, beta, mix, stats, flag, gscale: torch.Tensor) -> torch.Tensor:
    B, N = logits.shape
    d = torch.empty_like(logits)
    L.check(L.lib().msu_loss_bwd(L.dt(logits), logits.data_ptr(), target.data_ptr(), B, N, alpha, beta, mix,
                                 stats.data_ptr(), flag.data_ptr(), gscale.data_ptr(), d.data_ptr(), L.stream_ptr()),
            "msu_loss_bwd")
    return d


def metrics(inp: torch.Tensor, label_or_gt: torch.Tensor, pred_bin: Optional[torch.Tensor], from_logits: bool,
            thr: float, want_pred: bool = False, prob_f32: bool = False):
    """inp [B, N]; returns (counts int64 [B,4] = tp,fp,fn,tn, soft float64 [B,8], pred|None).
    from_logits: pred = sigmoid(inp) rounded to inp's dtype before the threshold (torch.sigmoid semantics on that dtype), or —
    `prob_f32` — kept in fp32 (pred comes back as fp32)."""
    B, N = inp.shape
    dev = inp.device
    wc = torch.empty(B * 64 * 4, dtype=torch.int64, device=dev)
    wsf = torch.empty(B * 64 * 8, dtype=torch.float64, device=dev)
    counts = torch.empty(B, 4, dtype=torch.int64, device=dev)
    soft = torch.empty(B, 8, dtype=torch.float64, device=dev)
    f32p = bool(from_logits and prob_f32)
    pred = torch.empty(inp.shape, dtype=torch.float32 if f32p else inp.dtype, device=dev) if (want_pred and from_logits) else None
    L.check(L.lib().msu_metrics(L.dt(inp), (2 if f32p else 1) if from_logits else 0, inp.data_ptr(), label_or_gt.data_ptr(),
                                L.ptr(pred_bin), B, N, thr, wc.data_ptr(), wsf.data_ptr(), counts.data_ptr(),
                                soft.data_ptr(), L.ptr(pred), L.stream_ptr()), "msu_metrics")
    return counts, soft, pred
