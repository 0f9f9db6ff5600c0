"""Batch-sharded data parallelism for the MS-UNet hot path: one process per GPU, gradients averaged with
NCCL over NVLink 5 / NVSwitch in flat buckets that are launched on a communication stream as soon as their
last gradient landed, so the exchange overlaps the rest of backward.

Replaces the reference's single-process `nn.DataParallel` (trainer.py:96-97), which re-broadcasts all
weights every step and reduces every gradient onto GPU 0.  The network has only LayerNorms and the loss is
a mean of per-sample losses, so sharding the batch is exact: averaging the per-rank gradients (equal shard
sizes) equals the gradient of the global-batch loss.

* Bucket order is learnt, not guessed: the first backward records the order in which gradients become ready
  (the shared concat_back_dim weights land late, the two dead decoder stacks never do — SURVEY.md facts 2, 5)
  and buckets are cut along that order.
* Gradients are WRITTEN into the buckets by the kernels that produce them: every parameter's bucket slice is
  registered as its gradient slot (`ops.register_grad_slots`), the backward Functions allocate their dW / db
  outputs there, and autograd adopts the slice as `.grad`.  Only gradients that autograd had to sum (shared
  weights) or accumulate are copied.
* Every parameter gets the ready hook as soon as it requires grad, so parameters unfrozen after wrapping (the reference's staged
  `unfreeze_encoder`, trainer.py:253-287) are reduced too: they take the slow path for one step and the
  buckets are re-learnt on the next.
* `ShardedAdamW` replaces all-reduce + a replicated optimizer step by reduce-scatter ->
  `msu_adamw_step` on this rank's shard -> all-gather of the updated parameters, per bucket, on the
  communication stream, i.e. the optimizer step is overlapped with backward as well.

Works under CUDA-graph capture (the comm stream forks from and joins the capturing stream through events).
"""
from __future__ import annotations

import contextlib
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn


class _Bucket:
    __slots__ = ("params", "offsets", "flat", "pending", "numel", "views", "shard")

    def __init__(self, params: List[nn.Parameter], device, dtype, multiple: int = 4):
        self.params = params
        self.offsets = []
        n = 0
        for p in params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4  # keep slices 16-byte aligned
        self.numel = (n + multiple - 1) // multiple * multiple
        self.flat = torch.zeros(self.numel, dtype=dtype, device=device)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, params)]
        self.pending = len(params)
        self.shard = None           # ShardedAdamW state of this bucket


class DataParallelB200(nn.Module):
    def __init__(self, module: nn.Module, bucket_mb: float = 32.0, process_group=None, grad_slots: bool = True):
        super().__init__()
        self.module = module
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.grad_slots = grad_slots
        self._params = list(module.parameters())    # every parameter gets the ready hook as soon as it requires grad
        self._order: List[nn.Parameter] = []        # gradient-ready order seen in the first backward
        self._buckets: Optional[List[_Bucket]] = None
        self._where = {}                            # param -> (bucket, index)
        self._late: List[nn.Parameter] = []
        self._comm_stream = None
        self._sync = True
        self._sharded = None                        # ShardedAdamW attached to this wrapper
        self.stats = {"direct": 0, "copied": 0, "late": 0, "recut": 0}
        self._backend = dist.get_backend(process_group) if dist.is_initialized() else "none"
        self._hooked = set()
        self._hook_trainable()
        if self.world > 1:
            self._broadcast_parameters()

    # -- plumbing -------------------------------------------------------------------------------
    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("module"), name)

    def _hook_trainable(self):
        """The ready hook can only be registered on tensors that require grad: parameters frozen at construction get theirs at
        the first forward after they were unfrozen (checked every forward: one pass over ~440 flags)."""
        if len(self._hooked) == len(self._params):
            return
        for p in self._params:
            if p.requires_grad and id(p) not in self._hooked:
                p.register_post_accumulate_grad_hook(self._on_grad_ready)
                self._hooked.add(id(p))

    def forward(self, *a, **kw):
        self._hook_trainable()
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

    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation over micro-batches: backward passes inside this context only accumulate into `.grad`
        (in place, i.e. into the bucket slices); the first backward outside it exchanges the sums."""
        prev, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = prev

    # -- gradient hooks -------------------------------------------------------------------------
    def _on_grad_ready(self, p: nn.Parameter):
        if self.world == 1 or not self._sync:
            return
        if self._buckets is None:
            self._order.append(p)
            return
        loc = self._where.get(p)
        if loc is None:  # no gradient when the buckets were cut (frozen then, unfrozen now): reduce it at the end of this step
            self._late.append(p)
            return
        b, i = loc
        if p.grad.data_ptr() != b.views[i].data_ptr():      # not produced in place (summed / accumulated by autograd): copy
            b.views[i].copy_(p.grad)
            self.stats["copied"] += 1
        else:
            self.stats["direct"] += 1
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b: _Bucket):
        s = self._stream(b.flat.device)
        if s is not None:
            s.wait_stream(torch.cuda.current_stream(b.flat.device))
            with torch.cuda.stream(s):
                self._exchange(b)
        else:
            self._exchange(b)

    def _exchange(self, b: _Bucket):
        if self._sharded is not None:
            self._sharded._bucket_step(b)
        else:
            self._all_reduce(b.flat)

    def _all_reduce(self, t: torch.Tensor):
        if self._backend == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.pg)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
            t.div_(self.world)

    def _cut_buckets(self):
        from . import ops
        order = [p for p in self._order if p.grad is not None]
        self._buckets, cur, size = [], [], 0
        self._where = {}
        mult = 4 * self.world
        seen = set()
        for p in order:
            if id(p) in seen:
                continue
            seen.add(id(p))
            cur.append(p)
            size += p.numel() * 4
            if size >= self.bucket_bytes:
                self._buckets.append(_Bucket(cur, p.device, torch.float32, mult))
                cur, size = [], 0
        if cur:
            self._buckets.append(_Bucket(cur, cur[0].device, torch.float32, mult))
        for b in self._buckets:
            for i, p in enumerate(b.params):
                self._where[p] = (b, i)
        self._late = []
        if self.grad_slots and self._buckets and self._buckets[0].flat.is_cuda:
            ops.register_grad_slots({p: v for b in self._buckets for p, v in zip(b.params, b.views)})

    def bucket_summary(self):
        return [] if not self._buckets else [(len(b.params), b.numel * 4) for b in self._buckets]

    def finish_gradient_sync(self):
        """Call after backward(): joins the comm stream and points every .grad at its averaged bucket slice."""
        if self.world == 1:
            return
        from . import ops
        if self._buckets is None:
            # first step: no overlap yet — reduce everything now, then cut buckets along the observed order
            for p in self._order:
                if p.grad is not None:
                    self._all_reduce(p.grad)
            self._cut_buckets()
            self._order = []
            ops.next_grad_pass()
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
                if p.grad is not None and p.grad.data_ptr() != b.views[i].data_ptr():
                    p.grad = b.views[i]
            b.pending = len(b.params)
        if self._late and self._sharded is not None:
            raise RuntimeError("parameters became trainable after ShardedAdamW was built: re-create the optimizer after unfreezing "
                               "(their shard state does not exist)")
        if self._late:
            # parameters that became trainable after the buckets were cut: reduce them one by one now and re-learn the
            # bucket layout during the next backward
            for p in self._late:
                self._all_reduce(p.grad)
            self.stats["late"] += len(self._late)
            self.stats["recut"] += 1
            self._late = []
            if self._sharded is None:
                ops.clear_grad_slots([p for b in self._buckets for p in b.params])
                self._buckets, self._where, self._order = None, {}, []
        ops.next_grad_pass()


# ------------------------------------------------------------------------------------------------
# Sharded optimizer step fused with the gradient exchange
# ------------------------------------------------------------------------------------------------
class ShardedAdamW:
    """AdamW whose state and update are sharded over the data-parallel ranks and whose collectives replace the gradient
    all-reduce: per bucket, as soon as its last gradient landed,

        reduce-scatter(AVG) of the flat gradient bucket  ->  `msu_adamw_step` on this rank's 1/N of the bucket
        ->  all-gather of the updated parameter shard into the flat parameter bucket,

    all on the communication stream, overlapped with the rest of backward.  Parameters become views of flat fp32 buckets
    (same values, same names); exp_avg / exp_avg_sq exist only for the local shard.  Arithmetic is torch.optim.AdamW's
    (trainer.py:143-152: decay / no-decay groups by name), element for element — the update of an element does not depend on
    which rank performs it, so the result equals all-reduce + `FusedAdamW` bit for bit (tests/test_dp_gloo.py,
    tests/test_gpu_optim.py).  `step()` only joins; hyper-parameters are read from `param_groups` when a step's first bucket is
    launched, so `lr` schedulers that run between steps (trainer.py:321-322) are honoured.  No GradScaler: this path computes
    in bf16 with fp32 gradients and needs no loss scaling (a scaler's skip decision would need every gradient before the first
    update, which is what the overlap removes)."""

    def __init__(self, dp: DataParallelB200, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if dp._buckets is None:
            raise RuntimeError("ShardedAdamW needs the bucket layout: run one forward/backward + finish_gradient_sync() first")
        if dp._sharded is not None:
            raise RuntimeError("this DataParallelB200 already has a sharded optimizer")
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.param_groups = [{**self.defaults, **g, "params": list(g["params"])} for g in groups]
        self.dp = dp
        self.world, self.rank = dp.world, dp.rank
        self._gidx = {p: gi for gi, g in enumerate(self.param_groups) for p in g["params"]}
        self.step_count = 0
        self._prepared = False
        self._chunk = None
        with torch.no_grad():
            for b in dp._buckets:
                self._adopt(b)
        if dp.grad_slots and dp._buckets and dp._buckets[0].flat.is_cuda:      # the parameters moved: re-key their gradient slots
            from . import ops
            ops.register_grad_slots({p: v for b in dp._buckets for p, v in zip(b.params, b.views)})
        dp._sharded = self

    # -- layout ---------------------------------------------------------------------------------
    def _adopt(self, b: _Bucket):
        """Flat parameter bucket (parameters re-pointed at its slices), local shard range and its (parameter, range) segments."""
        dev = b.flat.device
        S = b.numel // self.world
        lo, hi = self.rank * S, (self.rank + 1) * S
        pflat = torch.zeros(b.numel, dtype=torch.float32, device=dev)
        segs = []
        for p, off in zip(b.params, b.offsets):
            n = p.numel()
            pflat[off:off + n].copy_(p.detach().reshape(-1))
            p.data = pflat[off:off + n].view_as(p)
            a, e = max(off, lo), min(off + n, hi)
            if a < e and p in self._gidx:
                segs.append((p, a, e - a))                    # element range [a, a+len) of the bucket, inside this rank's shard
        b.shard = dict(S=S, lo=lo, pflat=pflat, gshard=torch.zeros(S, dtype=torch.float32, device=dev),
                       m=torch.zeros(S, dtype=torch.float32, device=dev), v=torch.zeros(S, dtype=torch.float32, device=dev),
                       segs=segs, table=None)

    def _table(self, b: _Bucket):
        """Device table of this bucket's segments in the record layout of csrc/optim.cu (MsuAdamTensor) + block maps."""
        from . import _lib as L
        from .optim import _REC_DTYPE
        sh = b.shard
        if self._chunk is None:
            self._chunk = int(L.lib().msu_adamw_chunk()) if b.flat.is_cuda else 8192
        n = len(sh["segs"])
        tab = np.zeros(n, dtype=_REC_DTYPE)
        bt, bc = [], []
        for i, (p, a, ln) in enumerate(sh["segs"]):
            loc = a - sh["lo"]
            tab["p"][i] = sh["pflat"].data_ptr() + 4 * a
            tab["g"][i] = sh["gshard"].data_ptr() + 4 * loc
            tab["m"][i] = sh["m"].data_ptr() + 4 * loc
            tab["v"][i] = sh["v"].data_ptr() + 4 * loc
            tab["n"][i] = ln
            k = (ln + self._chunk - 1) // self._chunk
            bt.append(np.full(k, i, dtype=np.int32))
            bc.append(np.arange(k, dtype=np.int32))
        dev = b.flat.device
        sh["table"] = tab
        sh["gidx"] = np.array([self._gidx[p] for p, _, _ in sh["segs"]], dtype=np.int64)
        sh["bt"] = torch.from_numpy(np.concatenate(bt) if bt else np.zeros(0, np.int32)).to(dev)
        sh["bc"] = torch.from_numpy(np.concatenate(bc) if bc else np.zeros(0, np.int32)).to(dev)
        # two pinned staging buffers: a buffer is rewritten only after the copy that read it two steps ago has completed
        sh["host"] = [torch.empty(max(n, 1) * 64, dtype=torch.uint8) for _ in range(2)]
        if dev.type == "cuda":
            sh["host"] = [h.pin_memory() for h in sh["host"]]
        sh["ev"], sh["flip"] = [None, None], 0
        sh["devt"] = torch.empty(max(n, 1) * 64, dtype=torch.uint8, device=dev)

    def prepare_step(self):
        """Upload the per-segment coefficients of the NEXT update (current `param_groups`, step count + 1) on the current
        stream.  Called lazily when a step's first bucket is launched; call it explicitly before replaying a captured step."""
        t = float(self.step_count + 1)
        g = self.param_groups
        lr = np.array([float(x["lr"]) for x in g]); wd = np.array([float(x["weight_decay"]) for x in g])
        b1 = np.array([float(x["betas"][0]) for x in g]); b2 = np.array([float(x["betas"][1]) for x in g])
        eps = np.array([float(x["eps"]) for x in g])
        for b in self.dp._buckets:
            sh = b.shard
            if sh["table"] is None:
                self._table(b)
            tab, gi = sh["table"], sh["gidx"]
            if len(gi):
                tab["decay"] = 1.0 - lr[gi] * wd[gi]
                tab["step_size"] = lr[gi] / (1.0 - b1[gi] ** t)
                tab["inv_bias2_sqrt"] = 1.0 / np.sqrt(1.0 - b2[gi] ** t)
                tab["beta1"], tab["beta2"], tab["eps"] = b1[gi], b2[gi], eps[gi]
                k = sh["flip"]
                sh["flip"] ^= 1
                if sh["ev"][k] is not None:
                    sh["ev"][k].synchronize()
                sh["host"][k].numpy()[:len(gi) * 64] = tab.view(np.uint8).reshape(-1)
                sh["devt"].copy_(sh["host"][k], non_blocking=True)
                if sh["devt"].is_cuda:
                    sh["ev"][k] = torch.cuda.Event()
                    sh["ev"][k].record()
        self._prepared = True

    def after_replay(self):
        """Bookkeeping of one replayed (CUDA-graph) step whose capture contained `step()`: `prepare_step(); graph.replay();
        after_replay()`."""
        self._count()

    # -- per-bucket exchange + update (runs on the communication stream) ------------------------------
    def _bucket_step(self, b: _Bucket):
        dp, sh = self.dp, b.shard
        if not self._prepared:
            if b.flat.is_cuda and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("ShardedAdamW.prepare_step() must be called before capturing / replaying a step")
            self.prepare_step()
        lo, S = sh["lo"], sh["S"]
        if dp._backend == "nccl":
            dist.reduce_scatter_tensor(sh["gshard"], b.flat, op=dist.ReduceOp.AVG, group=dp.pg)
            self._apply(b)
            dist.all_gather_into_tensor(sh["pflat"], sh["pflat"][lo:lo + S], group=dp.pg)
        else:
            # portable path (gloo, used by the host-logic tests on CPU and by the single-GPU two-process test): the same data
            # movement expressed with all-reduce only, the one collective every backend implements for every device
            dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=dp.pg)
            torch.div(b.flat[lo:lo + S], self.world, out=sh["gshard"])
            self._apply(b)
            tmp = torch.zeros_like(sh["pflat"])
            tmp[lo:lo + S].copy_(sh["pflat"][lo:lo + S])
            dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=dp.pg)
            sh["pflat"].copy_(tmp)

    def _apply(self, b: _Bucket):
        """One `msu_adamw_step` launch over this bucket's shard segments (CUDA only: there is no CPU fallback)."""
        from . import _lib as L
        sh = b.shard
        if not b.flat.is_cuda:
            raise RuntimeError("ShardedAdamW updates parameters with the CUDA kernel msu_adamw_step: CPU tensors are not supported")
        nblk = int(sh["bt"].numel())
        if nblk:
            L.check(L.lib().msu_adamw_step(sh["devt"].data_ptr(), sh["bt"].data_ptr(), sh["bc"].data_ptr(), nblk, None, None,
                                           L.stream_ptr()), "msu_adamw_step")

    # -- optimizer protocol ---------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    def step(self, closure=None):
        """The update already ran bucket by bucket during backward; this joins it (finish_gradient_sync) and counts the step."""
        if closure is not None:
            raise NotImplementedError("closures are not supported by the overlapped sharded step")
        self.dp.finish_gradient_sync()
        dev = self.dp._buckets[0].flat.device
        if not (dev.type == "cuda" and torch.cuda.is_current_stream_capturing()):
            self._count()            # a captured step is counted per replay (after_replay)

    def _count(self):
        self.step_count += 1
        self._prepared = False
        plist = [p for b in self.dp._buckets for p in b.params]
        torch.autograd.graph.increment_version(plist)       # parameters were written through raw pointers / collectives

    def state_dict(self):
        """torch.optim.AdamW layout (`step`, `exp_avg`, `exp_avg_sq` per parameter, full size) gathered from the shards."""
        state, order = {}, [p for g in self.param_groups for p in g["params"]]
        index = {p: i for i, p in enumerate(order)}
        for b in self.dp._buckets:
            sh = b.shard
            full = {}
            for key in ("m", "v"):
                out = torch.zeros(b.numel, dtype=torch.float32, device=b.flat.device)
                if self.dp._backend == "nccl":
                    dist.all_gather_into_tensor(out, sh[key], group=self.dp.pg)
                else:
                    out[sh["lo"]:sh["lo"] + sh["S"]].copy_(sh[key])
                    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.dp.pg)
                full[key] = out
            for p, off in zip(b.params, b.offsets):
                if p in index:
                    state[index[p]] = {"step": torch.tensor(float(self.step_count)),
                                       "exp_avg": full["m"][off:off + p.numel()].view_as(p).clone(),
                                       "exp_avg_sq": full["v"][off:off + p.numel()].view_as(p).clone()}
        groups = [{**{k: v for k, v in g.items() if k != "params"}, "params": [index[p] for p in g["params"]]}
                  for g in self.param_groups]
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        order = [p for g in self.param_groups for p in g["params"]]
        index = {p: i for i, p in enumerate(order)}
        for g, sg in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in sg.items() if k != "params"})
        steps = [float(s["step"]) for s in sd["state"].values()]
        self.step_count = int(max(steps)) if steps else 0
        with torch.no_grad():
            for b in self.dp._buckets:
                sh = b.shard
                for p, a, ln in sh["segs"]:
                    st = sd["state"].get(index[p])
                    if st is None:
                        continue
                    off = b.offsets[[q is p for q in b.params].index(True)]
                    loc = a - sh["lo"]
                    sh["m"][loc:loc + ln].copy_(st["exp_avg"].reshape(-1)[a - off:a - off + ln])
                    sh["v"][loc:loc + ln].copy_(st["exp_avg_sq"].reshape(-1)[a - off:a - off + ln])
        self._prepared = False


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
