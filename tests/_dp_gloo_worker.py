"""One gloo rank of tests/test_dp_gloo.py (launched as a subprocess: rank world port outfile)."""
import os
import sys

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


def main():
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
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
    c, s = all_gather_image_stats(torch.full((2, 4), rank, dtype=torch.int64),
                                  torch.full((2, 8), float(rank), dtype=torch.float64))
    if rank == 0:
        torch.save({"grads": grads, "sd": {k: v.clone() for k, v in m.state_dict().items()},
                    "buckets": m.bucket_summary(), "c": c, "s": s}, out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
