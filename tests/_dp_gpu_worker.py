"""One rank of tests/test_gpu_dp.py: two processes share cuda:0 and talk through gloo (NCCL refuses two ranks on one device), so
the REAL CUDA path — gradients written into bucket slices by the kernels, msu_adamw_step on a shard — runs on a one-GPU box.
Usage: rank world port outfile"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import msunet_oracle as O
    from semantic_segmentation_of_stylegan2_artifacts_b200 import dp as DP
    from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys
    from semantic_segmentation_of_stylegan2_artifacts_b200.optim import FusedAdamW
    dev = torch.device("cuda:0")
    cfg = O.Cfg(img_size=64, **O.T32)
    sd = O.make_weights(cfg)
    crit = DynamicLoss(alpha=0.2, beta=0.8, tversky_bce_mix=0.45)
    X, Y = O.make_inputs(cfg, 4, real_last=False)
    n = 4 // world
    xs, ys = X[rank * n:(rank + 1) * n].to(dev), Y[rank * n:(rank + 1) * n].to(dev)
    noise_all = O.draw_sd_noise(cfg, 4, 0.2, seed=11)
    mine = {k: (a[rank * n:(rank + 1) * n], b[rank * n:(rank + 1) * n]) for k, (a, b) in noise_all.items()}

    def build(prec):
        m = MSUNetSys(img_size=64, embed_dim=32, depths=[2, 2, 2, 2], num_heads=[1, 2, 4, 8], drop_path_rate=0.2)
        m.load_state_dict(sd, strict=True)
        return m.set_precision(prec).to(dev).train()

    def groups(m):
        named = [(k, p) for k, p in m.named_parameters() if not k.startswith(O.DEAD_PREFIXES)]
        nd = [p for k, p in named if p.ndim == 1 or k.endswith(".bias") or "norm" in k.lower()]
        dc = [p for k, p in named if not (p.ndim == 1 or k.endswith(".bias") or "norm" in k.lower())]
        return [{"params": dc, "weight_decay": 0.05}, {"params": nd, "weight_decay": 0.0}]

    res = {}
    for prec in ("fp32", "bf16"):
        # ---- (1) gradient exchange: 3 steps (learn order, bucketed, bucketed) on this rank's shard
        m = DP.DataParallelB200(build(prec), bucket_mb=0.25)
        m.module.inject_drop_path_noise(mine)
        for step in range(3):
            for p in m.parameters():
                p.grad = None
            loss = crit(m(xs), ys)
            loss.backward()
            m.finish_gradient_sync()
        torch.cuda.synchronize()
        res[prec] = {"grads": {k: p.grad.detach().cpu().clone() for k, p in m.module.named_parameters() if p.grad is not None},
                     "stats": dict(m.stats), "buckets": m.bucket_summary()}
        # ---- (2) sharded AdamW (on the same wrapper) vs replicated FusedAdamW on a second replica, 3 steps each
        ref = DP.DataParallelB200(build(prec), bucket_mb=0.25, grad_slots=False)
        ref.module.inject_drop_path_noise(mine)
        ropt = FusedAdamW(groups(ref.module), lr=1e-3, betas=(0.9, 0.95))
        sopt = DP.ShardedAdamW(m, groups(m.module), lr=1e-3, betas=(0.9, 0.95))
        for step in range(3):
            for mm, oo in ((ref, ropt), (m, sopt)):
                oo.zero_grad(set_to_none=True)
                crit(mm(xs), ys).backward()
                if oo is ropt:
                    mm.finish_gradient_sync()
                oo.step()
        torch.cuda.synchronize()
        res[prec]["sharded"] = {k: v.detach().cpu().clone() for k, v in m.module.state_dict().items()}
        res[prec]["replicated"] = {k: v.detach().cpu().clone() for k, v in ref.module.state_dict().items()}
        res[prec]["shard_state"] = sopt.state_dict()["state"][0]["exp_avg"].cpu()
        res[prec]["repl_state"] = ropt.state_dict()["state"][0]["exp_avg"].cpu()
        from semantic_segmentation_of_stylegan2_artifacts_b200 import ops
        ops.clear_grad_slots()
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
