"""`timm.scheduler.cosine_lr.CosineLRScheduler` restated for the arguments trainer.py:160-169 passes (single cycle, epoch
based, linear warm-up with `warmup_prefix`): lr(t) = warmup_lr_init + t (base - warmup_lr_init) / warmup_t for t < warmup_t, then
lr_min + (base - lr_min) (1 + cos(pi t' / t_initial)) / 2 with t' = t - warmup_t, lr_min after the cycle."""
import math


class CosineLRScheduler:
    def __init__(self, optimizer, t_initial, lr_min=0.0, warmup_t=0, warmup_lr_init=0.0, warmup_prefix=False, cycle_limit=0,
                 t_in_epochs=True, **kw):
        self.optimizer = optimizer
        self.t_initial, self.lr_min, self.warmup_t, self.warmup_lr_init = t_initial, lr_min, warmup_t, warmup_lr_init
        self.warmup_prefix, self.cycle_limit = warmup_prefix, cycle_limit
        for g in optimizer.param_groups:
            g.setdefault("initial_lr", g["lr"])
        self.base_values = [g["initial_lr"] for g in optimizer.param_groups]
        if warmup_t:
            self._set([warmup_lr_init] * len(self.base_values))

    def _set(self, values):
        for g, v in zip(self.optimizer.param_groups, values):
            g["lr"] = v

    def _get_lr(self, t):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * (b - self.warmup_lr_init) / self.warmup_t for b in self.base_values]
        if self.warmup_prefix:
            t = t - self.warmup_t
        if self.cycle_limit and t >= self.t_initial * self.cycle_limit:
            return [self.lr_min for _ in self.base_values]
        tc = t % self.t_initial
        return [self.lr_min + 0.5 * (b - self.lr_min) * (1 + math.cos(math.pi * tc / self.t_initial)) for b in self.base_values]

    def step(self, epoch, metric=None):
        self._set(self._get_lr(epoch))

    def step_update(self, num_updates, metric=None):
        pass
