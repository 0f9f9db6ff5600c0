"""Host -> device input staging for the training loop.

The reference moves every batch with a blocking `.to(device)` on the compute stream (trainer.py:299-300), so the copy
(67 MB per 16 x 512^2 fp32 batch) is serialised with the step.  `CudaPrefetcher` wraps any iterable of batches (dicts / tuples /
tensors, as the reference's DataLoader yields them): batch i+1 is copied from pinned host memory on a dedicated copy stream while
step i runs; `next()` hands out device tensors after making the compute stream wait for that copy.  The device tensors are a ring
of persistent buffers: a batch stays valid until two further batches have been drawn.

With `stage_uint8=True` the loader yields the RAW arrays of dataset/dataset.py:41-42 — {'image': uint8 [B,H,W,3], 'label': uint8
[B,H,W]} and optionally 'flip': bool [B] — and the `/255`, CHW transpose, flip and `>127` of dataset.py:13-16, 49-63 run on the
device (`ops.stage_u8`, one launch on the compute stream when the batch is handed out): 4 bytes per pixel cross PCIe instead of 16,
and the batch that comes out is
bit-identical to what the reference's transform would have produced on the host."""
from __future__ import annotations

from typing import Any, Iterable, Iterator

import torch


def _to_device(obj: Any, device, keep: list, slot: dict, path: str = "") -> Any:
    """Copy every tensor of a batch structure into the persistent device buffer of its position (`slot`: path -> tensor; a buffer
    is (re)allocated when the shape or dtype at that position changes, e.g. a last short batch)."""
    if torch.is_tensor(obj):
        src = obj if obj.is_pinned() or obj.is_cuda else obj.pin_memory()
        keep.append(src)                      # the pinned source must outlive the asynchronous copy
        buf = slot.get(path)
        if buf is None or buf.shape != src.shape or buf.dtype != src.dtype:
            buf = slot[path] = torch.empty(src.shape, dtype=src.dtype, device=device)
        buf.copy_(src, non_blocking=True)
        return buf
    if isinstance(obj, dict):
        return {k: _to_device(v, device, keep, slot, f"{path}/{k}") for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(v, device, keep, slot, f"{path}/{i}") for i, v in enumerate(obj))
    return obj


class CudaPrefetcher:
    """Batches live in a ring of RING persistent device buffers per tensor position (no allocation and no `record_stream` per
    step: allocating 64 MB on a side stream every step makes the caching allocator fall back to cudaMalloc whenever the recycled
    blocks are still guarded, a device-wide stall).  A batch is valid until RING - 1 further batches have been drawn; the copy
    stream waits (stream-ordered, no host sync) for the compute stream to be past the consumer of the slot it overwrites."""
    RING = 3

    def __init__(self, loader: Iterable, device=None, stage_uint8: bool = False):
        self.loader = loader
        self.stage_uint8 = stage_uint8
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CudaPrefetcher stages batches for a CUDA device (there is no CPU path)")
        self.stream = torch.cuda.Stream(device=self.device)
        self._slots = [dict() for _ in range(self.RING)]
        self._consumed = [None] * self.RING       # event on the compute stream: everything enqueued while the slot was current

    def __len__(self):
        return len(self.loader)

    def __iter__(self) -> Iterator:
        it = iter(self.loader)
        k = 0
        nxt = self._stage(it, k)
        while nxt is not None:
            batch, ev, _keep = nxt
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            if self.stage_uint8:              # 15 us on the compute stream: its outputs live in that stream's allocator pool
                batch = self._stage_u8(batch)
            nxt = self._stage(it, (k + 1) % self.RING)     # start the next copy before the caller launches this step
            yield batch
            # the caller has enqueued its work on this batch: the slot may be overwritten once the compute stream is past it
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            self._consumed[k] = done
            k = (k + 1) % self.RING

    def _stage(self, it, k):
        try:
            host = next(it)
        except StopIteration:
            return None
        keep: list = []
        with torch.cuda.stream(self.stream):
            if self._consumed[k] is not None:
                self.stream.wait_event(self._consumed[k])
            batch = _to_device(host, self.device, keep, self._slots[k])
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return batch, ev, keep

    @staticmethod
    def _stage_u8(batch):
        from . import ops
        if not isinstance(batch, dict) or "image" not in batch:
            raise ValueError("stage_uint8 expects dict batches with a uint8 'image' [B,H,W,3] (and 'label' [B,H,W])")
        out = dict(batch)
        img, lab = ops.stage_u8(batch["image"], batch.get("label"), batch.get("flip"))
        out["image"] = img
        if lab is not None:
            out["label"] = lab
        out.pop("flip", None)
        return out
