"""One gloo rank of tests/test_dp_gloo.py (launched as a subprocess: rank world port outfile [scenario])."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Toy(nn.Module):
    """Mimics the MS-UNet traits that matter to the reducer: a weight used twice, a dead branch."""

    def __init__(self):
        super().__init__()
        self.a = nn.Linear(8, 16)
        self.shared = nn.Linear(16, 16)
        self.dead = nn.Linear(16, 16)
        self.out = nn.Linear(16, 1)
        self.norm = nn.LayerNorm(16)

    def forward(self, x):
        h = self.shared(torch.tanh(self.a(x)))
        _ = self.dead(h.detach())          # evaluated, never consumed
        h = self.norm(self.shared(torch.tanh(h)))
        return self.out(h).squeeze(-1)


def data():
    g = torch.Generator().manual_seed(7)
    return torch.randn(8, 8, generator=g), torch.randn(8, generator=g)


def adamw_table_restatement(table_bytes: np.ndarray, n: int):
    """The arithmetic of csrc/optim.cu:adamw_elem applied through the same device table (CPU stand-in for the CUDA kernel in the
    host-logic test: pointers are host addresses here)."""
    import ctypes
    from semantic_segmentation_of_stylegan2_artifacts_b200.optim import _REC_DTYPE
    tab = table_bytes[:n * 64].view(_REC_DTYPE)
    for r in tab:
        k = int(r["n"])
        arr = lambda ptr: np.ctypeslib.as_array((ctypes.c_float * k).from_address(int(ptr)))
        p, g, m, v = arr(r["p"]), arr(r["g"]), arr(r["m"]), arr(r["v"])
        f = np.float32
        p *= f(r["decay"])
        m += (g - m) * (f(1.0) - f(r["beta1"]))
        v[:] = v * f(r["beta2"]) + (f(1.0) - f(r["beta2"])) * g * g
        p -= f(r["step_size"]) * (m / (np.sqrt(v) * f(r["inv_bias2_sqrt"]) + f(r["eps"])))


def scenario_grads(rank, world, out):
    from semantic_segmentation_of_stylegan2_artifacts_b200.dp import DataParallelB200, all_gather_image_stats
    torch.manual_seed(100 + rank)          # different init per rank: the wrapper must broadcast rank 0's weights
    m = DataParallelB200(Toy(), bucket_mb=0.0005)
    X, Y = data()
    n = 8 // world
    xs, ys = X[rank * n:(rank + 1) * n], Y[rank * n:(rank + 1) * n]
    grads = []
    for step in range(3):                  # step 0 learns the order, steps 1-2 use the bucketed path
        for p in m.parameters():
            p.grad = None
        loss = ((m(xs) - ys) ** 2).mean()
        loss.backward()
        m.finish_gradient_sync()
        grads.append({k: (None if p.grad is None else p.grad.clone()) for k, p in m.module.named_parameters()})
    # gradient accumulation: two micro-batches per rank under no_sync == one pass over the rank's shard
    for p in m.parameters():
        p.grad = None
    h = n // 2
    with m.no_sync():
        (((m(xs[:h]) - ys[:h]) ** 2).mean() / 2).backward()
    (((m(xs[h:]) - ys[h:]) ** 2).mean() / 2).backward()
    m.finish_gradient_sync()
    grads.append({k: (None if p.grad is None else p.grad.clone()) for k, p in m.module.named_parameters()})
    c, s = all_gather_image_stats(torch.full((2, 4), rank, dtype=torch.int64),
                                  torch.full((2, 8), float(rank), dtype=torch.float64))
    if rank == 0:
        torch.save({"grads": grads, "sd": {k: v.clone() for k, v in m.state_dict().items()},
                    "buckets": m.bucket_summary(), "c": c, "s": s}, out)


def scenario_unfreeze(rank, world, out):
    """freeze -> wrap -> train -> unfreeze (trainer.py:253-287): the newly trainable parameters must be reduced."""
    from semantic_segmentation_of_stylegan2_artifacts_b200.dp import DataParallelB200
    torch.manual_seed(5)
    toy = Toy()
    for p in toy.a.parameters():
        p.requires_grad = False
    m = DataParallelB200(toy, bucket_mb=0.0005)
    opt = torch.optim.SGD(m.parameters(), lr=0.05)
    X, Y = data()
    n = 8 // world
    xs, ys = X[rank * n:(rank + 1) * n], Y[rank * n:(rank + 1) * n]
    log = []
    for step in range(6):
        if step == 3:
            for p in toy.a.parameters():
                p.requires_grad_(True)
        opt.zero_grad(set_to_none=True)
        ((m(xs) - ys) ** 2).mean().backward()
        m.finish_gradient_sync()
        opt.step()
        log.append({k: (None if p.grad is None else p.grad.clone()) for k, p in toy.named_parameters()})
    flat = torch.cat([p.detach().reshape(-1) for p in toy.parameters()])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    if rank == 0:
        torch.save({"params": both, "log": log, "stats": dict(m.stats), "buckets": m.bucket_summary()}, out)


def scenario_sharded(rank, world, out):
    """reduce-scatter -> shard update -> all-gather (dp.ShardedAdamW) vs torch.optim.AdamW on the global batch."""
    from semantic_segmentation_of_stylegan2_artifacts_b200 import dp as DP
    torch.manual_seed(9)
    toy = Toy()
    m = DP.DataParallelB200(toy, bucket_mb=0.0005)
    X, Y = data()
    n = 8 // world
    xs, ys = X[rank * n:(rank + 1) * n], Y[rank * n:(rank + 1) * n]
    init = {k: v.clone() for k, v in toy.state_dict().items()}
    ((m(xs) - ys) ** 2).mean().backward()          # learn the bucket layout
    m.finish_gradient_sync()
    decay = [p for k, p in toy.named_parameters() if p.dim() > 1 and not k.startswith("dead.")]
    no_decay = [p for k, p in toy.named_parameters() if p.dim() <= 1 and not k.startswith("dead.")]
    opt = DP.ShardedAdamW(m, [{"params": decay, "weight_decay": 0.1}, {"params": no_decay, "weight_decay": 0.0}],
                          lr=1e-2, betas=(0.9, 0.95), eps=1e-8)
    # CPU stand-in for the CUDA kernel (same table, same arithmetic); the product path raises on CPU tensors
    try:
        opt._apply(m._buckets[0])
        raised = False
    except RuntimeError:
        raised = True
    opt._apply = lambda b: adamw_table_restatement(b.shard["devt"].numpy(), len(b.shard["segs"]))
    losses = []
    for step in range(4):
        if step == 2:
            for g in opt.param_groups:
                g["lr"] = 5e-3                      # a scheduler between steps (trainer.py:321-322)
        opt.zero_grad()
        loss = ((m(xs) - ys) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    sd = opt.state_dict()
    flat = torch.cat([p.detach().reshape(-1) for p in toy.parameters()])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    if rank == 0:
        torch.save({"init": init, "final": {k: v.clone() for k, v in toy.state_dict().items()}, "ranks": both, "raised": raised,
                    "opt_state": sd, "shards": [(b.shard["S"], len(b.shard["segs"])) for b in m._buckets]}, out)


def main():
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    scenario = sys.argv[5] if len(sys.argv) > 5 else "grads"
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    {"grads": scenario_grads, "unfreeze": scenario_unfreeze, "sharded": scenario_sharded}[scenario](rank, world, out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
