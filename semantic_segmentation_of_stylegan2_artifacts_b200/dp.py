"""Batch-sharded data parallelism for the MS-UNet hot path: one process per GPU, gradients averaged with
NCCL all-reduce over NVLink 5 / NVSwitch in flat buckets that are launched on a communication stream as
soon as their last gradient lands, so the exchange overlaps the rest of backward.

Replaces the reference's single-process `nn.DataParallel` (trainer.py:96-97), which re-broadcasts all
weights every step and reduces every gradient onto GPU 0.  The network has only LayerNorms and the loss is
a mean of per-sample losses, so sharding the batch is exact: averaging the per-rank gradients (equal shard
sizes) equals the gradient of the global-batch loss.

Bucket order is learnt, not guessed: the first backward records the order in which gradients become ready
(the shared concat_back_dim weights land late, the two dead decoder stacks never do — SURVEY.md facts 2, 5)
and buckets are cut along that order.  Works under CUDA-graph capture (the comm stream forks from and joins
the capturing stream through events).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn


class _Bucket:
    __slots__ = ("params", "offsets", "flat", "pending", "numel", "work")

    def __init__(self, params: List[nn.Parameter], device, dtype):
        self.params = params
        self.offsets = []
        n = 0
        for p in params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4  # keep slices 16-byte aligned
        self.numel = n
        self.flat = torch.zeros(n, dtype=dtype, device=device)
        self.pending = len(params)
        self.work = None


class DataParallelB200(nn.Module):
    def __init__(self, module: nn.Module, bucket_mb: float = 32.0, process_group=None):
        super().__init__()
        self.module = module
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self._params = [p for p in module.parameters() if p.requires_grad]
        self._order: List[nn.Parameter] = []        # gradient-ready order seen in the first backward
        self._buckets: Optional[List[_Bucket]] = None
        self._where = {}                            # param -> (bucket, index)
        self._comm_stream = None
        self._backend = dist.get_backend(process_group) if dist.is_initialized() else "none"
        for p in self._params:
            p.register_post_accumulate_grad_hook(self._on_grad_ready)
        if self.world > 1:
            self._broadcast_parameters()

    # -- plumbing -------------------------------------------------------------------------------
    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("module"), name)

    def forward(self, *a, **kw):
        return self.module(*a, **kw)

    def state_dict(self, *a, **kw):  # checkpoints keep the reference's key names (no "module." prefix)
        return self.module.state_dict(*a, **kw)

    def load_state_dict(self, *a, **kw):
        return self.module.load_state_dict(*a, **kw)

    def _broadcast_parameters(self):
        with torch.no_grad():
            for t in list(self.module.parameters()) + list(self.module.buffers()):
                dist.broadcast(t, src=0, group=self.pg)

    def _stream(self, device):
        if device.type != "cuda":
            return None
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=device)
        return self._comm_stream

    # -- gradient hooks -------------------------------------------------------------------------
    def _on_grad_ready(self, p: nn.Parameter):
        if self.world == 1:
            return
        if self._buckets is None:
            self._order.append(p)
            return
        loc = self._where.get(p)
        if loc is None:  # a parameter that had no gradient when the buckets were cut: reduce it at the end
            self._late.append(p)
            return
        b, i = loc
        n = p.numel()
        b.flat[b.offsets[i]:b.offsets[i] + n].copy_(p.grad.reshape(-1))
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b: _Bucket):
        s = self._stream(b.flat.device)
        if s is not None:
            s.wait_stream(torch.cuda.current_stream(b.flat.device))
            with torch.cuda.stream(s):
                self._all_reduce(b.flat)
        else:
            self._all_reduce(b.flat)

    def _all_reduce(self, t: torch.Tensor):
        if self._backend == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.pg)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
            t.div_(self.world)

    def _cut_buckets(self):
        order = [p for p in self._order if p.grad is not None]
        self._buckets, cur, size = [], [], 0
        for p in order:
            cur.append(p)
            size += p.numel() * 4
            if size >= self.bucket_bytes:
                self._buckets.append(_Bucket(cur, p.device, torch.float32))
                cur, size = [], 0
        if cur:
            self._buckets.append(_Bucket(cur, cur[0].device, torch.float32))
        for b in self._buckets:
            for i, p in enumerate(b.params):
                self._where[p] = (b, i)
        self._late = []

    def bucket_summary(self):
        return [] if not self._buckets else [(len(b.params), b.numel * 4) for b in self._buckets]

    def finish_gradient_sync(self):
        """Call after backward(): joins the comm stream and points every .grad at its averaged bucket slice."""
        if self.world == 1:
            return
        if self._buckets is None:
            # first step: no overlap yet — reduce everything now, then cut buckets along the observed order
            for p in self._order:
                if p.grad is not None:
                    self._all_reduce(p.grad)
            self._cut_buckets()
            self._order = []
            return
        dev = self._buckets[0].flat.device
        for b in self._buckets:
            if b.pending != 0 and b.pending != len(b.params):
                # some gradients of this bucket never arrived this step (e.g. frozen encoder): send what we have
                self._launch(b)
        s = self._stream(dev)
        if s is not None:
            torch.cuda.current_stream(dev).wait_stream(s)
        for b in self._buckets:
            if b.pending == len(b.params):
                continue  # nothing arrived: parameters keep grad None
            for i, p in enumerate(b.params):
                if p.grad is not None:
                    p.grad = b.flat[b.offsets[i]:b.offsets[i] + p.numel()].view_as(p)
            b.pending = len(b.params)
        for p in self._late:
            self._all_reduce(p.grad)
        self._late = []


def all_gather_image_stats(counts: torch.Tensor, soft: torch.Tensor, group=None):
    """Sharded validation: every rank counts its own images; per-image records (4 int64 + 8 float64) are
    all-gathered so that rank 0 can aggregate exactly like scripts/validation_functions.py:148-211."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return counts, soft
    w = dist.get_world_size(group)
    oc = torch.empty((w * counts.shape[0],) + tuple(counts.shape[1:]), dtype=counts.dtype, device=counts.device)
    os_ = torch.empty((w * soft.shape[0],) + tuple(soft.shape[1:]), dtype=soft.dtype, device=soft.device)
    dist.all_gather_into_tensor(oc, counts.contiguous(), group=group)
    dist.all_gather_into_tensor(os_, soft.contiguous(), group=group)
    return oc, os_
