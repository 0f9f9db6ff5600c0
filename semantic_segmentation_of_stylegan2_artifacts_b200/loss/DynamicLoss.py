"""`DynamicLoss(roi_thresh, alpha, beta, tversky_bce_mix)(output, target)` — the reference's
loss/DynamicLoss.py:73-111 as one fused CUDA forward and one fused backward, with no host
synchronisation (the reference performs 1+3B of them and a Python loop over the batch)."""
from __future__ import annotations

import torch

from ..functional import DynamicLossFn


class DynamicLoss(torch.nn.Module):
    def __init__(self, roi_thresh=0.04, alpha=0.4, beta=0.6, tversky_bce_mix=0.5):
        super().__init__()
        self.roi_thresh = roi_thresh  # accepted and unused, as in the reference (:75)
        self.alpha, self.beta, self.tversky_bce_mix = float(alpha), float(beta), float(tversky_bce_mix)

    def forward(self, output, target):
        if target.dim() == 3:  # (B,H,W) -> (B,1,H,W)
            target = target.unsqueeze(1)
        B, B_t = output.size(0), target.size(0)
        if B != B_t:
            raise ValueError(f"Batchsize from ouptut {B} not equal to batchsize target {B_t}")
        if output[0].numel() != target[0].numel():
            raise ValueError(f"target shape {tuple(target.shape)} does not match output {tuple(output.shape)}")
        return DynamicLossFn.apply(output, target, self.alpha, self.beta, self.tversky_bce_mix)

    @torch.no_grad()
    def per_sample(self, output, target):
        """Per-image losses [B] (fp32, device-resident, no host sync): what `forward` returns for each image alone — used by the
        batched validation loop (the reference evaluates one image per call, validation_functions.py:89-104), so the
        {0,255} label decision of :87-88 is taken per image, not per batch."""
        from .. import ops
        if target.dim() == 3:
            target = target.unsqueeze(1)
        B = output.size(0)
        if B != target.size(0):
            raise ValueError(f"Batchsize from ouptut {B} not equal to batchsize target {target.size(0)}")
        ops._need_cuda(output, "logits")
        lg = output.contiguous().view(B, -1)
        tg = target.float().contiguous().view(B, -1)
        return ops.loss_per_sample(lg, tg, self.alpha, self.beta, self.tversky_bce_mix)
