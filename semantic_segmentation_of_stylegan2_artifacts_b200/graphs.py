"""Whole-step CUDA graph for the training loop (trainer.py:296-318 with fixed batch shapes).

An eager MS-UNet step is ~800 kernel launches from Python: on a B200 the host, not the GPU, sets the pace (26.1 ms eager vs 24.6 ms
replayed for the T96 512x512 batch-16 iteration).  `GraphedStep` captures forward + criterion + backward once and replays it per
batch; the optimizer step stays outside (eager), exactly where trainer.py:315 has it:

    step = GraphedStep(model, criterion, example_image, example_label)      # after model.train(), before the loop
    for batch in loader:
        loss = step(batch["image"], batch["label"])     # gradients are in p.grad when this returns (stream-ordered)
        optimizer.step()

Everything random inside the step (stochastic depth, attention dropout) draws from PyTorch's CUDA generator, which CUDA graphs
advance per replay, so replays see fresh noise.  The weight shadows are re-derived inside the graph (functional.refresh_shadows),
so optimizer updates between replays are picked up.  Pure plumbing: no kernels of its own."""
from __future__ import annotations

from typing import Callable, Optional

import torch


class GraphedStep:
    def __init__(self, model: torch.nn.Module, criterion: Callable, image: torch.Tensor, label: torch.Tensor, warmup: int = 3,
                 post_backward: Optional[Callable[[], None]] = None):
        if not image.is_cuda:
            raise RuntimeError("GraphedStep captures a CUDA graph: the example batch must live on the GPU (no CPU path)")
        self.model, self.criterion = model, criterion
        self.image, self.label = image.clone(), label.clone()          # static input buffers the graph reads
        self.params = [p for p in model.parameters() if p.requires_grad]
        self._post = post_backward

        side = torch.cuda.Stream(device=image.device)
        side.wait_stream(torch.cuda.current_stream(image.device))
        with torch.cuda.stream(side):                                  # warm-up off the legacy stream (allocations, shadows, plans)
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream(image.device).wait_stream(side)
        torch.cuda.synchronize(image.device)
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()
        self.grads = [p.grad for p in self.params]                     # static gradient buffers owned by the graph's pool

    def _eager(self) -> torch.Tensor:
        for p in self.params:
            p.grad = None
        loss = self.criterion(self.model(self.image), self.label)
        loss.backward()
        if self._post is not None:
            self._post()
        return loss.detach()

    def __call__(self, image: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        if image.shape != self.image.shape or label.shape != self.label.shape:
            raise ValueError(f"GraphedStep was captured for {tuple(self.image.shape)} / {tuple(self.label.shape)}, "
                             f"got {tuple(image.shape)} / {tuple(label.shape)} (drop_last=True keeps the batch shape fixed)")
        self.image.copy_(image, non_blocking=True)
        self.label.copy_(label, non_blocking=True)
        self.graph.replay()
        for p, g in zip(self.params, self.grads):                      # survives optimizer.zero_grad(set_to_none=True)
            p.grad = g
        return self.loss
