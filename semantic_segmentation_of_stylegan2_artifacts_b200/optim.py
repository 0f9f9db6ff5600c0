"""Fused multi-tensor AdamW for the MS-UNet training loop.

`FusedAdamW` keeps torch.optim.AdamW's constructor, param-group handling, `state_dict()` layout (`step`, `exp_avg`,
`exp_avg_sq`) and arithmetic (reference: trainer.py:143-152 builds `optim.AdamW([{decay}, {no_decay}], lr, betas, eps,
amsgrad=False)`, trainer.py:315 steps it through a GradScaler), but one `msu_adamw_step` launch updates every parameter:
16 B read + 12 B written per parameter instead of PyTorch's per-tensor / per-op passes.  CUDA fp32 parameters only — anything
else raises (no CPU fallback).  `patch_torch_adamw()` makes `torch.optim.AdamW` resolve to it so that the reference's
trainer runs unchanged.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L


class _AdamRec(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64),
                ("decay", C.c_float), ("step_size", C.c_float), ("inv_bias2_sqrt", C.c_float), ("beta1", C.c_float),
                ("beta2", C.c_float), ("eps", C.c_float)]


_REC_DTYPE = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8"), ("decay", "<f4"),
                       ("step_size", "<f4"), ("inv_bias2_sqrt", "<f4"), ("beta1", "<f4"), ("beta2", "<f4"), ("eps", "<f4")])
assert _REC_DTYPE.itemsize == C.sizeof(_AdamRec) == 64


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, **unused):
        if amsgrad:
            raise NotImplementedError("amsgrad is not used by the reference (trainer.py:151) and not implemented")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or not 0.0 <= weight_decay:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False))
        self._layout = None      # cached table / block maps for the current set of parameters with gradients
        self._chunk = None
        # torch.amp.GradScaler protocol (trainer.py:182, 314-316): the scaler hands its scale and the inf flag over as
        # `self.grad_scale` / `self.found_inf` instead of unscaling every gradient in a separate pass; the kernel multiplies
        # by 1/scale while it reads the gradient
        self._step_supports_amp_scaling = True

    # ------------------------------------------------------------------
    def _build(self):
        """(Re)build the cached layout: the parameters that currently have gradients, their static table columns and
        the block -> (tensor, chunk) maps.  Rebuilt only when that set changes."""
        if self._chunk is None:
            self._chunk = int(L.lib().msu_adamw_chunk())
        plist, gidx = [], []
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("FusedAdamW handles CUDA fp32 parameters with dense fp32 gradients only (no CPU fallback)")
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdamW needs contiguous parameters")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                plist.append(p)
                gidx.append(gi)
        if not plist:
            self._layout = None
            return
        dev = plist[0].device
        n = len(plist)
        table = np.zeros(n, dtype=_REC_DTYPE)
        table["p"] = [p.data_ptr() for p in plist]
        table["m"] = [self.state[p]["exp_avg"].data_ptr() for p in plist]
        table["v"] = [self.state[p]["exp_avg_sq"].data_ptr() for p in plist]
        table["n"] = [p.numel() for p in plist]
        steps = np.array([float(self.state[p]["step"]) for p in plist], dtype=np.float64)
        bt, bc = [], []
        for i, p in enumerate(plist):
            k = (p.numel() + self._chunk - 1) // self._chunk
            bt.append(np.full(k, i, dtype=np.int32))
            bc.append(np.arange(k, dtype=np.int32))
        self._layout = dict(plist=plist, ids=[id(p) for p in plist], gidx=np.array(gidx), table=table, steps=steps, dev=dev,
                            bt=torch.from_numpy(np.concatenate(bt)).to(dev), bc=torch.from_numpy(np.concatenate(bc)).to(dev),
                            host=[torch.empty(n * 64, dtype=torch.uint8).pin_memory() for _ in range(2)],
                            devt=[torch.empty(n * 64, dtype=torch.uint8, device=dev) for _ in range(2)],
                            ev=[None, None], flip=0,
                            ptrs=(table["p"].copy(), table["m"].copy(), table["v"].copy()))

    def _sync_steps(self):
        """Write the host-side step counters into the `step` state tensors (state_dict layout of torch.optim.AdamW)."""
        lay = self._layout
        if lay is not None:
            for p, t in zip(lay["plist"], lay["steps"]):
                self.state[p]["step"].fill_(float(t))

    def state_dict(self):
        self._sync_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._layout = None

    def add_param_group(self, param_group):
        if getattr(self, "_layout", None) is not None:
            self._sync_steps()
        super().add_param_group(param_group)
        self._layout = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale, found_inf = getattr(self, "grad_scale", None), getattr(self, "found_inf", None)
        if found_inf is not None and float(found_inf.item()) != 0.0:
            return loss          # overflow: skip the step and its step count, as GradScaler._maybe_opt_step does for stock optimizers
        inv_scale = None if grad_scale is None else (1.0 / grad_scale.detach().double()).float().reshape(1)
        lay = self._layout
        live = [p for group in self.param_groups for p in group["params"] if p.grad is not None]
        if lay is None or len(live) != len(lay["ids"]) or any(id(p) != i for p, i in zip(live, lay["ids"])):
            if lay is not None:
                self._sync_steps()
            self._build()
            lay = self._layout
            if lay is None:
                return loss
        else:   # storages can move (e.g. .to(), load_state_dict): cheap pointer check
            tp = np.fromiter((p.data_ptr() for p in lay["plist"]), dtype=np.uint64, count=len(live))
            if not np.array_equal(tp, lay["ptrs"][0]):
                self._sync_steps()
                self._build()
                lay = self._layout
        plist, table = lay["plist"], lay["table"]
        grads = []
        for p in plist:
            g = p.grad
            if g.dtype != torch.float32 or g.is_sparse:
                raise RuntimeError("FusedAdamW needs dense fp32 gradients")
            grads.append(g if g.is_contiguous() else g.contiguous())
        table["g"] = np.fromiter((g.data_ptr() for g in grads), dtype=np.uint64, count=len(grads))
        lay["steps"] += 1.0
        t = lay["steps"]
        gi = lay["gidx"]
        lr = np.array([float(g["lr"]) for g in self.param_groups])[gi]
        wd = np.array([float(g["weight_decay"]) for g in self.param_groups])[gi]
        b1 = np.array([float(g["betas"][0]) for g in self.param_groups])[gi]
        b2 = np.array([float(g["betas"][1]) for g in self.param_groups])[gi]
        eps = np.array([float(g["eps"]) for g in self.param_groups])[gi]
        table["decay"] = 1.0 - lr * wd
        table["step_size"] = lr / (1.0 - b1 ** t)
        table["inv_bias2_sqrt"] = 1.0 / np.sqrt(1.0 - b2 ** t)
        table["beta1"], table["beta2"], table["eps"] = b1, b2, eps
        k = lay["flip"]
        lay["flip"] ^= 1
        if lay["ev"][k] is not None:
            lay["ev"][k].synchronize()          # the copy that used this pinned buffer two steps ago has completed
        lay["host"][k].numpy()[:] = table.view(np.uint8).reshape(-1)
        lay["devt"][k].copy_(lay["host"][k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        lay["ev"][k] = ev
        L.check(L.lib().msu_adamw_step(lay["devt"][k].data_ptr(), lay["bt"].data_ptr(), lay["bc"].data_ptr(),
                                       int(lay["bt"].numel()), L.ptr(inv_scale), None, L.stream_ptr()), "msu_adamw_step")
        # the kernel wrote through raw pointers: bump the autograd version counters so that everything keyed on them
        # (the bf16 weight shadows of functional.shadow, saved-tensor checks) sees the update
        torch.autograd.graph.increment_version(plist)
        self._keep = grads                      # contiguous gradient copies stay alive until the next step
        return loss


def patch_torch_adamw() -> None:
    """Make `torch.optim.AdamW` (what the reference's trainer.py constructs) resolve to FusedAdamW."""
    torch.optim.AdamW = FusedAdamW
