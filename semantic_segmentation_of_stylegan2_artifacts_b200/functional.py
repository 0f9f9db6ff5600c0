"""autograd.Function layer: each Function is one fused module of the MS-UNet with a hand-written
forward AND backward made only of C-ABI kernel launches (ops.py).  torch.autograd is used for graph
bookkeeping (accumulating gradients of shared weights, skip connections) and nothing else.

Activations are token-major `[rows, C]` tensors in the compute dtype (bf16 or fp32); parameters
stay fp32 `nn.Parameter`s and gradients are returned in fp32.

bf16 mode feeds the tcgen05 GEMMs: weights are read through cached bf16 shadows (plain for the forward,
transposed for dgrad so that every tensor-core operand is K-major), and gradient rows that autograd
hands over in pixel order are materialised once in window / depth-to-space order (msu_gather_rows) so
that dgrad and wgrad both read dense TMA tiles.  fp32 mode (the <=1e-3 parity mode) reads the fp32
parameters directly through the mapped-operand SIMT engine.
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops
from .ops import MAP_CONV3, MAP_MERGE, MAP_SHUFFLE, MAP_UNSHUFFLE, MAP_WINDOW, epilogue, gemm, operand

WS = 7
BF16 = torch.bfloat16


def window_geo(H: int, W: int, shift: int):
    """{H, W, Ph, Pw, sh, sw}; the shift is disabled when the padded map is one window
    (TV:models/swin_transformer.py:158-163)."""
    Ph = WS * ((H + WS - 1) // WS)
    Pw = WS * ((W + WS - 1) // WS)
    return [H, W, Ph, Pw, 0 if WS >= Ph else shift, 0 if WS >= Pw else shift]


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _al(*ts):
    """Parameters as the kernels want them: contiguous and 16-byte aligned (they are read with 128-bit loads).  nn.DataParallel
    replicas (trainer.py:96-97) are 4-byte-aligned slices of one broadcast buffer: those get an aligned private copy (the
    gradients still flow to the original tensors: backward returns them by position)."""
    out = tuple(t if (t is None or (t.is_contiguous() and t.data_ptr() % 16 == 0)) else t.contiguous().clone() for t in ts)
    return out if len(out) > 1 else out[0]


# ----------------------------------------------------------------------------------------------
# Weight-gradient side stream: dW / db GEMMs (and their split-K reduces) are off the critical path of a block's
# backward, so they are enqueued on a second stream and overlap the dgrad chain (small reduce kernels fill the SMs a
# persistent GEMM leaves idle).  Every Function joins the side stream before it returns, so tensor lifetimes stay
# those of the main stream (the caching allocator never sees a cross-stream use after free) and CUDA-graph capture
# records a plain fork / join.  MSUNET_B200_WGRAD_STREAM=0 keeps everything on one stream.
# ----------------------------------------------------------------------------------------------
import os as _os

_WG_ON = _os.environ.get("MSUNET_B200_WGRAD_STREAM", "1") != "0"
_STORE_GELU_GRAD = _os.environ.get("MSUNET_B200_STORE_GELU_GRAD", "1") != "0"   # MLP forward saves GELU'(h) instead of h
_FUSED_HEAD_LN = _os.environ.get("MSUNET_B200_FUSED_HEAD_LN", "1") != "0"   # head LayerNorm + 1x1 conv in the second conv's epilogue
_wg_streams: dict = {}


class _WgradFork:
    def __init__(self, dev):
        self.on = _WG_ON and ops.PROF is None      # per-op profiling keeps one stream (events bracket each launch)
        self.keep = []                             # temporaries read on the side stream: released after the join
        if self.on:
            st = _wg_streams.get(dev.index)
            if st is None:
                st = _wg_streams[dev.index] = torch.cuda.Stream(device=dev)
            self.side = st
            self.main = torch.cuda.current_stream(dev)

    def run(self, fn):
        """fn() enqueues GEMMs whose inputs are complete on the main stream at this point."""
        if not self.on:
            return fn()
        self.side.wait_stream(self.main)
        with torch.cuda.stream(self.side):
            fn()

    def join(self):
        if self.on:
            self.main.wait_stream(self.side)
        self.keep.clear()

    def __enter__(self):
        self._prev = ops.set_side(self if self.on else None)
        return self

    def __exit__(self, *exc):
        ops.set_side(self._prev)
        self.join()
        return False


# ----------------------------------------------------------------------------------------------
# bf16 weight shadows (caller-owned tensors, refreshed when the fp32 master changes)
# ----------------------------------------------------------------------------------------------
_shadows: dict = {}        # (id(param), mode, dtype) -> [weakref(param), version tag, shadow tensor, (R, Cc)]
_shadow_gen = [0]          # bumped whenever an entry is (re)allocated or dropped: refresh plans are rebuilt lazily
_cap_token = [None]        # identity of the most recent in-capture refresh (see refresh_shadows)


def _drop_shadow(key):
    _shadows.pop(key, None)
    _shadow_gen[0] += 1


def shadow(p: torch.Tensor, mode: int, R: int, Cc: int, shape, dtype=BF16) -> torch.Tensor:
    """prep_weight(mode) of parameter `p`, cached on (parameter identity, version, storage).  While a CUDA graph is being captured
    an entry only counts as fresh if `refresh_shadows` re-derived it inside this capture (the master may change between replays)."""
    key = (id(p), mode, dtype)
    ent = _shadows.get(key)
    if ent is not None and ent[0]() is p:
        if torch.cuda.is_current_stream_capturing():
            if ent[1] == ("captured", _cap_token[0]):
                return ent[2]
        elif ent[1] == (p._version, p.data_ptr()):
            return ent[2]
    t = ops.prep_weight(mode, p, R, Cc, shape, dtype)
    try:
        ref = weakref.ref(p, lambda _r, k=key: _drop_shadow(k))
    except TypeError:  # pragma: no cover
        ref = (lambda q=p: q)
    # a shadow made during capture lives in the graph's pool and is re-derived by every replay: never valid for eager use
    ver = ("captured", None) if torch.cuda.is_current_stream_capturing() else (p._version, p.data_ptr())
    _shadows[key] = [ref, ver, t, (R, Cc)]
    _shadow_gen[0] += 1
    return t


def _shadow_plan(params, state: dict):
    """Job table of every registered shadow of `params` (device-resident, rebuilt only when the registry changes)."""
    import numpy as np
    from . import _lib as L
    plan = state.get("plan")
    if plan is not None and plan["gen"] == _shadow_gen[0]:
        return plan
    rec = np.dtype([("src", "<u8"), ("dst", "<u8"), ("dst_t", "<u8"), ("R", "<i8"), ("C", "<i8"), ("mode", "<i4"), ("dtype", "<i4")])
    jobs, ents, plist = [], [], []
    for p in params:
        mine = {}
        for mode in (0, 1, 2, 3, 5):
            for dt in (BF16, torch.float32):
                e = _shadows.get((id(p), mode, dt))
                if e is not None and e[0]() is p and e[2].device == p.device:
                    mine[(mode, dt)] = e
        if not mine:
            continue
        plist.append(p)
        for dt in (BF16, torch.float32):
            e0, e1 = mine.get((0, dt)), mine.get((1, dt))
            if e0 is not None or e1 is not None:
                R, Cc = (e0 or e1)[3]
                jobs.append((p.data_ptr(), e0[2].data_ptr() if e0 else 0, e1[2].data_ptr() if e1 else 0, R, Cc, 0, L.dt(e0[2] if e0 else e1[2])))
                ents += [(e, p) for e in (e0, e1) if e is not None]
            for mode in (2, 3, 5):
                e = mine.get((mode, dt))
                if e is not None:
                    jobs.append((p.data_ptr(), e[2].data_ptr(), 0, e[3][0], e[3][1], mode, L.dt(e[2])))
                    ents.append((e, p))
    plan = {"gen": _shadow_gen[0], "ents": ents, "plist": plist, "ptrs": [p.data_ptr() for p in plist], "n_blocks": 0}
    if jobs:
        dev = plist[0].device
        table = np.array(jobs, dtype=rec)
        bj, bt = [], []
        for i, j in enumerate(jobs):
            nb = int(L.lib().msu_shadow_blocks(int(j[5]), int(j[3]), int(j[4])))
            if nb <= 0:
                raise RuntimeError("msu_shadow_blocks rejected a weight shadow job")
            bj.append(np.full(nb, i, dtype=np.int32))
            bt.append(np.arange(nb, dtype=np.int32))
        plan.update(table=torch.from_numpy(table.view(np.uint8).reshape(-1).copy()).to(dev),
                    bj=torch.from_numpy(np.concatenate(bj)).to(dev), bt=torch.from_numpy(np.concatenate(bt)).to(dev),
                    n_blocks=int(sum(len(x) for x in bj)))
    state["plan"] = plan
    return plan


def refresh_shadows(params, state: dict, force: bool = False) -> None:
    """Re-derive every registered shadow of `params` in ONE launch (`msu_refresh_shadows`) if any master changed since the shadows
    were made — after an optimizer step that is all of them, which the per-tensor path would redo as ~230 small launches.  During
    CUDA-graph capture the refresh is always recorded (replays must see the current masters) and marks the entries it covered as
    fresh for this capture.  Shadows that are not registered yet (first forward) are made lazily by `shadow()` as before.
    `state` is a caller-owned dict (one per model) holding the cached job table.  `force` refreshes without consulting the
    version counters: fused optimizers (`torch.optim.AdamW(fused=True)`, anything writing through raw pointers) update parameters
    without bumping them, so a training forward never trusts them."""
    from . import _lib as L
    capturing = torch.cuda.is_current_stream_capturing()
    plan = state.get("plan")
    fresh_plan = plan is not None and plan["gen"] == _shadow_gen[0] and all(p.data_ptr() == q for p, q in zip(plan["plist"], plan["ptrs"]))
    if not fresh_plan:
        if capturing:
            return            # building the table needs host -> device copies: leave this capture to the per-tensor path
        state.pop("plan", None)
        plan = _shadow_plan(params, state)
    if plan["n_blocks"] == 0:
        return
    if not capturing and not force and all(e[1] == (p._version, p.data_ptr()) for e, p in plan["ents"]):
        return
    L.check(L.lib().msu_refresh_shadows(plan["table"].data_ptr(), plan["bj"].data_ptr(), plan["bt"].data_ptr(), plan["n_blocks"],
                                        L.stream_ptr()), "msu_refresh_shadows")
    if capturing:
        _cap_token[0] = tok = object()
        for e, _p in plan["ents"]:
            e[1] = ("captured", tok)
    else:
        for e, p in plan["ents"]:
            e[1] = (p._version, p.data_ptr())


def w_fwd(w: torch.Tensor, dt):
    """B operand of Y = X W^T for a [N, K] weight."""
    if dt == BF16:
        N, K = w.shape
        return operand(shadow(w, 0, N, K, (N, K)))
    return operand(w)


def w_dgrad(w: torch.Tensor, dt, col0: int = 0, ncols: Optional[int] = None):
    """B operand of dX = dY W[:, col0:col0+ncols] for a [N, K] weight."""
    N, K = w.shape
    if dt == BF16:
        wt = shadow(w, 1, N, K, (K, N))          # W^T, K-major for the tensor core
        return operand(wt, offset=col0 * N)
    return operand(w, ld=K, orient=1, offset=col0)


def rows(t: torch.Tensor, M: int, N: int, dt, **kw):
    """A operand whose logical rows are read through a map / per-sample scale.  bf16: materialise once
    (dense rows for TMA) and return (operand, orient-1 operand); fp32: mapped operands, no copy."""
    mapped = kw.get("map", 0) != 0 or kw.get("rowscale") is not None
    if dt == BF16 and mapped:
        d = ops.gather_rows(operand(t, **kw), M, N, t)
        return operand(d), operand(d, orient=1)
    return operand(t, **kw), operand(t, orient=1, **kw)


# ----------------------------------------------------------------------------------------------
class SwinBlockFn(Function):
    """x + SD(attn(LN1 x)) then + SD(mlp(LN2 .)) — TV:models/swin_transformer.py:401-455."""

    @staticmethod
    def forward(ctx, x, n1w, n1b, qkvw, qkvb, projw, projb, table, n2w, n2b, f1w, f1b, f2w, f2b, sd1, sd2,
                B, H, W, nH, shift, attn_p=0.0, drop_seed=None):
        ops._need_cuda(x, "x")
        x = _c(x)
        n1w, n1b, qkvw, qkvb, projw, projb, table, n2w, n2b, f1w, f1b, f2w, f2b = _al(
            n1w, n1b, qkvw, qkvb, projw, projb, table, n2w, n2b, f1w, f1b, f2w, f2b)
        dev, dt = x.device, x.dtype
        Cd = x.shape[-1]
        T = B * H * W
        geo = window_geo(H, W, shift)
        nW = (geo[2] // WS) * (geo[3] // WS)
        Tw = B * nW * 49
        hid = f1w.shape[0]
        HW = H * W
        # LN1 + zero-pad + roll + window partition in one gather pass
        xw, mean1, rstd1 = ops.ln_fwd(x, n1w, n1b, Tw, Cd, out_map=MAP_WINDOW, geo=geo, n_stat_rows=T)
        qkv = torch.empty(Tw, 3 * Cd, dtype=dt, device=dev)
        gemm(operand(xw), w_fwd(qkvw, dt), epilogue(qkv, bias=qkvb), Tw, 3 * Cd, Cd, dev)
        bias = ops.relbias_expand(table, nH)
        need_grad = any(ctx.needs_input_grad)     # inference: nothing is kept for a backward (no lse, no GELU' tensor)
        if need_grad:
            o, lse = ops.winattn_fwd(qkv, bias, B * nW, nH, geo, attn_p, drop_seed, want_lse=True)
        else:
            o, lse = ops.winattn_fwd(qkv, bias, B * nW, nH, geo, attn_p, drop_seed), None
        # proj + window reverse + un-roll + crop + stochastic depth + residual in the GEMM epilogue
        x1 = torch.empty(T, Cd, dtype=dt, device=dev)
        gemm(operand(o), w_fwd(projw, dt),
             epilogue(x1, bias=projb, R=x, map=MAP_WINDOW, geo=geo, rowscale=sd1, rps=HW), Tw, Cd, Cd, dev)
        xn, mean2, rstd2 = ops.ln_fwd(x1, n2w, n2b, T, Cd)
        h = torch.empty(T, hid, dtype=dt, device=dev) if need_grad else None
        a = torch.empty(T, hid, dtype=dt, device=dev)
        # `h` holds GELU'(pre-activation), not the pre-activation (act = 2): the backward's dh epilogue is then a plain multiply
        # (act = 3) instead of re-deriving GELU' (15 of its 31 instructions per element; MSUNET_B200_STORE_GELU_GRAD=0: keep h)
        gemm(operand(xn), w_fwd(f1w, dt), epilogue(a, Cpre=h, bias=f1b, act=2 if (_STORE_GELU_GRAD and need_grad) else 1), T, hid, Cd, dev)
        x2 = torch.empty(T, Cd, dtype=dt, device=dev)
        gemm(operand(a), w_fwd(f2w, dt), epilogue(x2, bias=f2b, R=x1, rowscale=sd2, rps=HW), T, Cd, hid, dev)
        if need_grad:
            ctx.save_for_backward(x, n1w, n1b, qkvw, projw, n2w, n2b, f1w, f2w, sd1, sd2,
                                  xw, mean1, rstd1, qkv, bias, o, x1, xn, mean2, rstd2, h, a, qkvb, projb, table, f1b, f2b, lse)
        ctx.cfg = (B, H, W, nH, geo, nW)
        ctx.drop = (attn_p, drop_seed)
        return x2.view(B, H, W, Cd)

    @staticmethod
    @once_differentiable
    def backward(ctx, dx2):
        (x, n1w, n1b, qkvw, projw, n2w, n2b, f1w, f2w, sd1, sd2,
         xw, mean1, rstd1, qkv, bias, o, x1, xn, mean2, rstd2, h, a, qkvb, projb, table, f1b, f2b, lse) = ctx.saved_tensors
        B, H, W, nH, geo, nW = ctx.cfg
        dev, dt = x.device, x.dtype
        Cd = x.shape[-1]
        T, HW, Tw, hid = B * H * W, H * W, B * nW * 49, f1w.shape[0]
        dx2 = _c(dx2).view(T, Cd)
        f32 = dict(dtype=torch.float32, device=dev)
        with _WgradFork(dev) as wg:
            # ---- MLP half
            # stochastic-depth scale of the incoming gradient rows: large maps never materialise sd2 * dx2 — the dgrad GEMM scales its
            # output rows in the epilogue and the weight-gradient GEMM keeps every split-K slab inside one sample and scales whole
            # partials in its reduce (MsuOperand.rowscale); small maps (few tokens per sample) gather the scaled rows once
            fused_sd = dt == BF16 and sd2 is not None and HW >= 4096 and HW % 64 == 0
            if fused_sd:
                dy2, dy2t = operand(dx2), operand(dx2, orient=1, rowscale=sd2, rps=HW)
            else:
                dy2, dy2t = rows(dx2, T, Cd, dt, rowscale=sd2, rps=HW)
            # bias gradients ride along with the weight-gradient GEMMs (MsuEpilogue.colsum)
            db2, dW2 = ops.grad_out(f2b), ops.grad_out(f2w)
            wg.run(lambda: gemm(dy2t, operand(a, orient=1), epilogue(dW2, out_f32=True, colsum=db2), Cd, hid, T, dev))
            dh = torch.empty(T, hid, dtype=dt, device=dev)
            gemm(dy2, w_dgrad(f2w, dt),
                 epilogue(dh, H=h, ldh=hid, rowscale=sd2 if fused_sd else None, rps=HW if fused_sd else 0,
                          act=3 if _STORE_GELU_GRAD else 0), T, hid, Cd, dev)
            db1, dW1 = ops.grad_out(f1b), ops.grad_out(f1w)
            wg.run(lambda: gemm(operand(dh, orient=1), operand(xn, orient=1), epilogue(dW1, out_f32=True, colsum=db1), hid, Cd, T, dev))
            dxn = torch.empty(T, Cd, dtype=dt, device=dev)
            gemm(operand(dh), w_dgrad(f1w, dt), epilogue(dxn), T, Cd, hid, dev)
            # ---- attention half: gradient rows in window order (zero rows for the padding tokens).  bf16: the LayerNorm backward
            # that produces dx1 writes the scaled window-ordered copy itself (persistent buffer, padding rows zeroed once)
            if dt == BF16:
                dyw = ops.window_rows_buffer(Tw, Cd, geo, x)
                dx1, dn2w, dn2b = ops.ln_bwd_dual(dxn, x1, n2w, n2b, mean2, rstd2, T, Cd, dx2, dyw, geo, sd1, HW)
                dy1, dy1t = operand(dyw), operand(dyw, orient=1)
            else:
                dx1, dn2w, dn2b, _ = ops.ln_bwd(dxn, x1, n2w, n2b, mean2, rstd2, T, Cd, dres=dx2)
                dy1, dy1t = rows(dx1, Tw, Cd, dt, map=MAP_WINDOW, geo=geo, rowscale=sd1, rps=HW)
            dbp, dWp = ops.grad_out(projb), ops.grad_out(projw)
            wg.run(lambda: gemm(dy1t, operand(o, orient=1), epilogue(dWp, out_f32=True, colsum=dbp), Cd, Cd, Tw, dev))
            do = torch.empty(Tw, Cd, dtype=dt, device=dev)
            gemm(dy1, w_dgrad(projw, dt), epilogue(do), Tw, Cd, Cd, dev)
            dqkv, dtable = ops.winattn_bwd(qkv, bias, o, do, B * nW, nH, geo, *ctx.drop, table=table, lse=lse)
            dbqkv, dWqkv = ops.grad_out(qkvb), ops.grad_out(qkvw)
            wg.run(lambda: gemm(operand(dqkv, orient=1), operand(xw, orient=1), epilogue(dWqkv, out_f32=True, colsum=dbqkv),
                                3 * Cd, Cd, Tw, dev))
            dxw = torch.empty(Tw, Cd, dtype=dt, device=dev)
            gemm(operand(dqkv), w_dgrad(qkvw, dt), epilogue(dxw), Tw, Cd, 3 * Cd, dev)
            dx, dn1w, dn1b, _ = ops.ln_bwd(dxw, x, n1w, n1b, mean1, rstd1, T, Cd, dres=dx1, dy_map=MAP_WINDOW, geo=geo)
        return (dx.view(B, H, W, Cd), dn1w, dn1b, dWqkv, dbqkv, dWp, dbp, dtable, dn2w, dn2b, dW1, db1, dW2, db2,
                None, None, None, None, None, None, None, None, None)


# ----------------------------------------------------------------------------------------------
class PatchEmbedFn(Function):
    """Conv2d(3,E,4,4)/4 as patchify + GEMM, then LayerNorm(E) — network/model_parts.py:211-224."""

    @staticmethod
    def forward(ctx, img, pw, pb, nw, nb, dtype):
        ops._need_cuda(img, "image")
        pw, pb, nw, nb = _al(pw, pb, nw, nb)
        img = _c(img.float())
        dev = img.device
        B, _, S, _ = img.shape
        E = pw.shape[0]
        T = B * (S // 4) ** 2
        patches = ops.patchify4(img, dtype)
        w64 = shadow(pw, 5, E, 48, (E, 64), dtype)
        y = torch.empty(T, E, dtype=dtype, device=dev)
        gemm(operand(patches), operand(w64), epilogue(y, bias=pb), T, E, 64, dev)
        out, mean, rstd = ops.ln_fwd(y, nw, nb, T, E)
        ctx.save_for_backward(patches, y, nw, nb, mean, rstd, pw, pb)
        ctx.E = E
        return out.view(B, (S // 4) ** 2, E)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        patches, y, nw, nb, mean, rstd, pw, pb = ctx.saved_tensors
        E = ctx.E
        dev = y.device
        T = y.shape[0]
        dout = _c(dout).view(T, E)
        with _WgradFork(dev):
            dy, dnw, dnb, _ = ops.ln_bwd(dout, y, nw, nb, mean, rstd, T, E)
            dpb = ops.grad_out(pb)
            dW64 = torch.empty(E, 64, dtype=torch.float32, device=dev)
            gemm(operand(dy, orient=1), operand(patches, orient=1), epilogue(dW64, out_f32=True, colsum=dpb), E, 64, T, dev)
        dpw = ops.grad_out(pw, (E, 3, 4, 4))
        ops.prep_weight_into(6, dW64, dpw, E, 48)
        return None, dpw, dpb, dnw, dnb, None


class PatchMergeFn(Function):
    """2x2 gather + LayerNorm(4C) + Linear(4C,2C,no bias) — network/model_parts.py:87-95."""

    @staticmethod
    def forward(ctx, x, nw, nb, rw, B, H, W):
        x = _c(x)
        nw, nb, rw = _al(nw, nb, rw)
        dev, dt = x.device, x.dtype
        Cd = x.shape[-1]
        Tm = B * (H // 2) * (W // 2)
        geo = [H, W, Cd]
        xm, mean, rstd = ops.ln_fwd(x, nw, nb, Tm, 4 * Cd, in_map=MAP_MERGE, geo=geo)
        y = torch.empty(Tm, 2 * Cd, dtype=dt, device=dev)
        gemm(operand(xm), w_fwd(rw, dt), epilogue(y), Tm, 2 * Cd, 4 * Cd, dev)
        ctx.save_for_backward(x, nw, nb, rw, xm, mean, rstd)
        ctx.cfg = (B, H, W)
        return y.view(B, (H // 2) * (W // 2), 2 * Cd)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, nw, nb, rw, xm, mean, rstd = ctx.saved_tensors
        B, H, W = ctx.cfg
        dev, dt = x.device, x.dtype
        Cd = x.shape[-1]
        Tm = B * (H // 2) * (W // 2)
        dy = _c(dy).view(Tm, 2 * Cd)
        drw = ops.grad_out(rw)
        with _WgradFork(dev) as wg:
            wg.run(lambda: gemm(operand(dy, orient=1), operand(xm, orient=1), epilogue(drw, out_f32=True), 2 * Cd, 4 * Cd, Tm, dev))
            dxm = torch.empty(Tm, 4 * Cd, dtype=dt, device=dev)
            gemm(operand(dy), w_dgrad(rw, dt), epilogue(dxm), Tm, 4 * Cd, 2 * Cd, dev)
            dx, dnw, dnb, _ = ops.ln_bwd(dxm, x, nw, nb, mean, rstd, Tm, 4 * Cd, dx_map=MAP_MERGE, geo=[H, W, Cd])
        return dx, dnw, dnb, drw, None, None, None


class PatchExpandFn(Function):
    """Linear(C,2C,no bias) -> depth-to-space x2 (GEMM output map) -> LayerNorm(C/2) — model_parts.py:395-405."""

    @staticmethod
    def forward(ctx, x, ew, nw, nb, B, H, W):
        x = _c(x)
        ew, nw, nb = _al(ew, nw, nb)
        dev, dt = x.device, x.dtype
        Cd = x.shape[-1]
        T = B * H * W
        x2d = x.view(T, Cd)
        c2 = Cd // 2
        geo = [H, W, 2, c2]
        y = torch.empty(4 * T, c2, dtype=dt, device=dev)
        gemm(operand(x2d), w_fwd(ew, dt), epilogue(y, ldc=c2, map=MAP_SHUFFLE, geo=geo), T, 2 * Cd, Cd, dev)
        out, mean, rstd = ops.ln_fwd(y, nw, nb, 4 * T, c2)
        ctx.save_for_backward(x2d, ew, nw, nb, y, mean, rstd)
        ctx.cfg = (B, H, W, geo, x.shape)
        return out.view(B, 4 * H * W, c2)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x2d, ew, nw, nb, y, mean, rstd = ctx.saved_tensors
        B, H, W, geo, xshape = ctx.cfg
        dev, dt = y.device, y.dtype
        T, Cd = x2d.shape
        c2 = Cd // 2
        dout = _c(dout).view(4 * T, c2)
        # LayerNorm backward writes its result straight in the inverse depth-to-space layout [T, 2C]
        with _WgradFork(dev) as wg:
            dy, dnw, dnb, _ = ops.ln_bwd(dout, y, nw, nb, mean, rstd, 4 * T, c2, dx_map=MAP_UNSHUFFLE, geo=geo,
                                         dx_shape=(T, 2 * Cd))
            dew = ops.grad_out(ew)
            wg.run(lambda: gemm(operand(dy, orient=1), operand(x2d, orient=1), epilogue(dew, out_f32=True), 2 * Cd, Cd, T, dev))
            dx = torch.empty(T, Cd, dtype=dt, device=dev)
            gemm(operand(dy), w_dgrad(ew, dt), epilogue(dx), T, Cd, 2 * Cd, dev)
        return dx.view(xshape), dew, dnw, dnb, None, None, None


class ConcatLinearFn(Function):
    """Linear(2C,C)+bias on cat([x, skip], -1) without materialising the concat (dual-source K) —
    network/model_parts.py:792-793, 804-805, 823-824."""

    @staticmethod
    def forward(ctx, x, skip, w, b):
        x, skip = _c(x), _c(skip)
        w, b = _al(w, b)
        dev, dt = x.device, x.dtype
        Cd = x.shape[-1]
        T = x.numel() // Cd
        y = torch.empty(T, Cd, dtype=dt, device=dev)
        gemm(operand(x.view(T, Cd), t2=skip.view(T, Cd), ld2=Cd, k_split=Cd), w_fwd(w, dt), epilogue(y, bias=b),
             T, Cd, 2 * Cd, dev)
        ctx.save_for_backward(x, skip, w, b)
        return y.view(x.shape[0], -1, Cd)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, skip, w, b = ctx.saved_tensors
        dev, dt = x.device, x.dtype
        Cd = x.shape[-1]
        T = x.numel() // Cd
        dy = _c(dy).view(T, Cd)
        db, dw = ops.grad_out(b), ops.grad_out(w)

        def wgrads():
            gemm(operand(dy, orient=1), operand(x.view(T, Cd), orient=1), epilogue(dw, ldc=2 * Cd, out_f32=True, colsum=db),
                 Cd, Cd, T, dev)
            gemm(operand(dy, orient=1), operand(skip.view(T, Cd), orient=1),
                 epilogue(dw, ldc=2 * Cd, out_f32=True, offset=Cd), Cd, Cd, T, dev)

        with _WgradFork(dev) as wg:
            wg.run(wgrads)
            dx = torch.empty(T, Cd, dtype=dt, device=dev)
            dskip = torch.empty(T, Cd, dtype=dt, device=dev)
            gemm(operand(dy), w_dgrad(w, dt, 0, Cd), epilogue(dx), T, Cd, Cd, dev)
            gemm(operand(dy), w_dgrad(w, dt, Cd, Cd), epilogue(dskip), T, Cd, Cd, dev)
        return dx.view(x.shape), dskip.view(skip.shape), dw, db


class LayerNormFn(Function):
    """Plain LayerNorm over the last dim (norm / norm_up, network/model_parts.py:813, 827)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = _c(x)
        w, b = _al(w, b)
        Cd = x.shape[-1]
        nrows = x.numel() // Cd
        y, mean, rstd = ops.ln_fwd(x, w, b, nrows, Cd)
        ctx.save_for_backward(x, w, b, mean, rstd)
        return y.view(x.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w, b, mean, rstd = ctx.saved_tensors
        Cd = x.shape[-1]
        nrows = x.numel() // Cd
        dx, dw, db, _ = ops.ln_bwd(_c(dy), x, w, b, mean, rstd, nrows, Cd)
        return dx, dw, db


class HeadFn(Function):
    """FinalPatchExpand_X4_V2 + 1x1 output conv (network/model_parts.py:451-476, 842-846), NHWC end to end:
    Linear(E,16E)+GELU with the x4 depth-to-space as the GEMM output map, two implicit-GEMM 3x3 convs
    (+bias, GELU after the first), LayerNorm(E) fused with the 1x1 conv (a per-pixel dot product)."""

    @staticmethod
    def forward(ctx, x, ew, c1w, c1b, c2w, c2b, nw, nb, ow, B, r):
        x = _c(x)
        ew, c1w, c1b, c2w, c2b, nw, nb = _al(ew, c1w, c1b, c2w, c2b, nw, nb)
        dev, dt = x.device, x.dtype
        E = x.shape[-1]
        T = B * r * r
        S = 4 * r
        Mp = B * S * S
        x2d = x.view(T, E)
        sgeo = [r, r, 4, E]
        cgeo = [S, S, E]
        need_grad = any(ctx.needs_input_grad)     # inference: the GELU' / pre-activation tensors of the expand and the first conv are not written
        h0 = torch.empty(Mp, E, dtype=dt, device=dev) if need_grad else None
        a0 = torch.empty(Mp, E, dtype=dt, device=dev)
        # h0 / z1 hold GELU'(pre-activation) (act = 2), not the pre-activation: the two dgrad convolutions multiply by them (act = 3)
        # instead of deriving GELU' per element in their epilogues (3x3 conv dgrad 792 -> 705 us, forward 720 -> 739 us at 512^2 x 16)
        act_f = 2 if (_STORE_GELU_GRAD and need_grad) else 1
        gemm(operand(x2d), w_fwd(ew, dt), epilogue(a0, ldc=E, Cpre=h0, act=act_f, map=MAP_SHUFFLE, geo=sgeo), T, 16 * E, E, dev)
        wd = dt if dt == BF16 else torch.float32
        w1 = shadow(c1w, 2, E, E, (E, 9 * E), wd)
        w2 = shadow(c2w, 2, E, E, (E, 9 * E), wd)
        z1 = torch.empty(Mp, E, dtype=dt, device=dev) if need_grad else None
        a1 = torch.empty(Mp, E, dtype=dt, device=dev)
        gemm(operand(a0, ld=E, map=MAP_CONV3, geo=cgeo), operand(w1), epilogue(a1, Cpre=z1, bias=c1b, act=act_f),
             Mp, E, 9 * E, dev)
        owv = _al(_c(ow).view(E))
        if _FUSED_HEAD_LN and dt == BF16 and E % 32 == 0 and E <= 256 and S % 128 == 0:
            # LayerNorm + 1x1 conv ride in the second conv's epilogue (per-row statistics in the epilogue registers): the
            # [Mp, E] tensor is not read again, and not even written when no gradient is wanted (inference)
            z2 = torch.empty(Mp, E, dtype=dt, device=dev) if need_grad else None
            logits = torch.empty(Mp, dtype=dt, device=dev)
            mean = torch.empty(Mp, dtype=torch.float32, device=dev)
            rstd = torch.empty(2, Mp, dtype=torch.float32, device=dev)       # [rstd | dot statistic], see ops.ln_fwd
            gemm(operand(a1, ld=E, map=MAP_CONV3, geo=cgeo), operand(w2),
                 epilogue(z2, ldc=E, bias=c2b, lnd=(nw, nb, owv, logits, mean, rstd[0], rstd[1]), dtype=dt), Mp, E, 9 * E, dev)
            if z2 is None:
                z2 = logits          # placeholder: nothing of the head is needed without a backward
        else:
            z2 = torch.empty(Mp, E, dtype=dt, device=dev)
            gemm(operand(a1, ld=E, map=MAP_CONV3, geo=cgeo), operand(w2), epilogue(z2, bias=c2b), Mp, E, 9 * E, dev)
            logits, mean, rstd = ops.ln_fwd(z2, nw, nb, Mp, E, dotw=owv)
        if need_grad:
            ctx.save_for_backward(x2d, ew, c1w, c2w, nw, nb, owv, h0, a0, z1, a1, z2, mean, rstd, c1b, c2b)
        ctx.cfg = (B, r, x.shape, ow.shape)
        return logits.view(B, 1, S, S)

    @staticmethod
    @once_differentiable
    def backward(ctx, dlogits):
        x2d, ew, c1w, c2w, nw, nb, owv, h0, a0, z1, a1, z2, mean, rstd, c1b, c2b = ctx.saved_tensors
        B, r, xshape, owshape = ctx.cfg
        dev, dt = x2d.device, x2d.dtype
        T, E = x2d.shape
        S = 4 * r
        Mp = B * S * S
        sgeo = [r, r, 4, E]
        cgeo = [S, S, E]
        f32 = dict(dtype=torch.float32, device=dev)
        wd = dt if dt == BF16 else torch.float32
        dl = _c(dlogits).view(Mp)
        with _WgradFork(dev) as wg:
            dz2, dnw, dnb, dow = ops.ln_bwd(dl, z2, nw, nb, mean, rstd, Mp, E, dotw=owv)
            # conv2 (weight gradients on the side stream: they fill the tails of the dgrad convolutions)
            dc2b = ops.grad_out(c2b)
            dw2r = torch.empty(E, 9 * E, **f32)
            dc2w = ops.grad_out(c2w)

            def wgrad2():
                gemm(operand(dz2, orient=1), operand(a1, ld=E, orient=1, map=MAP_CONV3, geo=cgeo),
                     epilogue(dw2r, out_f32=True, colsum=dc2b), E, 9 * E, Mp, dev)
                ops.prep_weight_into(4, dw2r, dc2w, E, E)
            wg.run(wgrad2)
            w2f = shadow(c2w, 3, E, E, (E, 9 * E), wd)
            dz1 = torch.empty(Mp, E, dtype=dt, device=dev)
            act_b = 3 if _STORE_GELU_GRAD else 0
            gemm(operand(dz2, ld=E, map=MAP_CONV3, geo=cgeo), operand(w2f), epilogue(dz1, H=z1, ldh=E, act=act_b), Mp, E, 9 * E, dev)
            # conv1
            dc1b = ops.grad_out(c1b)
            dw1r = torch.empty(E, 9 * E, **f32)
            dc1w = ops.grad_out(c1w)

            def wgrad1():
                gemm(operand(dz1, orient=1), operand(a0, ld=E, orient=1, map=MAP_CONV3, geo=cgeo),
                     epilogue(dw1r, out_f32=True, colsum=dc1b), E, 9 * E, Mp, dev)
                ops.prep_weight_into(4, dw1r, dc1w, E, E)
            wg.run(wgrad1)
            w1f = shadow(c1w, 3, E, E, (E, 9 * E), wd)
            # conv1 dgrad (x gelu'(h0)) written by the GEMM epilogue in the inverse depth-to-space layout [T, 16E]
            dh0 = torch.empty(T, 16 * E, dtype=dt, device=dev)
            gemm(operand(dz1, ld=E, map=MAP_CONV3, geo=cgeo), operand(w1f),
                 epilogue(dh0, ldc=16 * E, H=h0, ldh=E, act=act_b, map=MAP_UNSHUFFLE, geo=sgeo), Mp, E, 9 * E, dev)
            dew = ops.grad_out(ew)
            wg.run(lambda: gemm(operand(dh0, orient=1), operand(x2d, orient=1), epilogue(dew, out_f32=True), 16 * E, E, T, dev))
            dx = torch.empty(T, E, dtype=dt, device=dev)
            gemm(operand(dh0), w_dgrad(ew, dt), epilogue(dx), T, E, 16 * E, dev)
        return (dx.view(xshape), dew, dc1w, dc1b, dc2w, dc2b, dnw, dnb, dow.view(owshape), None, None)


class DynamicLossFn(Function):
    """Fused per-sample BCE-with-logits + Tversky mix, batch mean — loss/DynamicLoss.py:82-111."""

    @staticmethod
    def forward(ctx, logits, target, alpha, beta, mix):
        ops._need_cuda(logits, "logits")
        B = logits.shape[0]
        lg = _c(logits).view(B, -1)
        tg = _c(target.float()).view(B, -1)
        loss, stats, flag = ops.loss_fwd(lg, tg, alpha, beta, mix)
        ctx.save_for_backward(lg, tg, stats, flag)
        ctx.cfg = (alpha, beta, mix, logits.shape)
        return loss.view(())

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        lg, tg, stats, flag = ctx.saved_tensors
        alpha, beta, mix, shape = ctx.cfg
        gs = _c(g.float()).view(1)
        d = ops.loss_bwd(lg, tg, alpha, beta, mix, stats, flag, gs)
        return d.view(shape), None, None, None, None


def drop_path_noise(p: float, training: bool, B: int, device) -> Optional[torch.Tensor]:
    """Row-mode stochastic depth noise Bernoulli(1-p)/(1-p) per sample (TV:ops/stochastic_depth.py:35-44)."""
    if not training or p == 0.0:
        return None
    keep = 1.0 - p
    n = torch.empty(B, dtype=torch.float32, device=device).bernoulli_(keep)
    if keep > 0.0:
        n.div_(keep)
    return n
