"""Whole training iteration as trainer.py:296-318 runs it (zero_grad, forward, DynamicLoss, backward, AdamW step) on the T96
512x512 batch-16 workload: img/s with the fused AdamW + single-launch weight-shadow refresh vs torch.optim.AdamW(fused=True).
Two forms: every launch eager, and fwd+loss+bwd replayed from a CUDA graph with the optimizer step issued eagerly after it
(its hyper-parameter table is uploaded every step).  Device-timed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import T96, synth_batch  # noqa: E402
import semantic_segmentation_of_stylegan2_artifacts_b200 as pkg  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.graphs import GraphedStep  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.optim import FusedAdamW  # noqa: E402

dev = torch.device("cuda:0")
B, S, steps = 16, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 10
x, y = synth_batch(B, S, 4321)
x, y = x.to(dev), y.to(dev)
crit = DynamicLoss(alpha=0.2, beta=0.8, tversky_bce_mix=0.45)
for name, mk in (("msu FusedAdamW", lambda ps: FusedAdamW(ps, lr=1e-4, weight_decay=0.01)),
                 ("torch AdamW fused=True", lambda ps: torch.optim.AdamW(ps, lr=1e-4, weight_decay=0.01, fused=True)),
                 ("no optimizer (fwd+bwd only)", None)):
    torch.manual_seed(1234)
    m = MSUNetSys(img_size=S, drop_path_rate=0.1, **T96).to(dev).train()
    opt = mk(m.parameters()) if mk else None

    def it():
        m.zero_grad(set_to_none=True)
        loss = crit(m(x), y)
        loss.backward()
        if opt is not None:
            opt.step()
        return loss

    for _ in range(3):
        it()
    torch.cuda.synchronize()
    for mode in ("eager", "GraphedStep + optimizer"):
        if mode != "eager":
            loss = None                       # drop the last eager autograd graph (its AccumulateGrad nodes pin the legacy stream)
            gstep = GraphedStep(m, crit, x, y, warmup=1)        # the public helper: fwd + loss + bwd in one CUDA graph

            def it():
                l = gstep(x, y)
                if opt is not None:
                    opt.step()
                return l

            for _ in range(2):
                it()
            torch.cuda.synchronize()
        n0 = pkg.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = it()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(f"{name:28s} [{mode:23s}]: {ms:7.2f} ms/iteration  {B / ms * 1e3:7.1f} img/s  loss {float(loss.detach()):.4f}", flush=True)
