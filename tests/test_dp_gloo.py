"""Host-side data-parallel logic on CPU: world_size-2 gloo ranks must reproduce the single-process gradient of
the concatenated batch (bucketing along the learnt ready order, shared weights, unused parameters)."""
import os
import socket
import subprocess
import sys

import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(tmp_path, scenario):
    out = str(tmp_path / f"{scenario}.pt")
    port = str(_free_port())
    worker = os.path.join(ROOT, "tests", "_dp_gloo_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", port, out, scenario]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=180) == 0
    return torch.load(out, weights_only=False)


def test_two_rank_gradients_equal_global_batch(tmp_path):
    from _dp_gloo_worker import Toy, data
    res = _run(tmp_path, "grads")
    grads, sd, buckets, c, s = res["grads"], res["sd"], res["buckets"], res["c"], res["s"]
    ref = Toy()
    ref.load_state_dict(sd)
    assert all(not k.startswith("module.") for k in sd)   # checkpoints keep the reference's key names
    X, Y = data()
    ((ref(X) - Y) ** 2).mean().backward()
    assert len(buckets) >= 2                      # the tiny bucket size forces several buckets
    assert len(grads) == 4                        # 3 plain steps + one accumulated over two micro-batches under no_sync()
    for step_grads in grads:
        for k, p in ref.named_parameters():
            if k.startswith("dead."):
                assert step_grads[k] is None and p.grad is None
            else:
                assert torch.allclose(step_grads[k], p.grad, rtol=1e-5, atol=1e-7), k
    assert c.shape == (4, 4) and c[:2].eq(0).all() and c[2:].eq(1).all()
    assert s.shape == (4, 8) and s[2:].eq(1.0).all()


def test_parameters_unfrozen_after_wrapping_are_reduced(tmp_path):
    """freeze_encoder -> wrap -> unfreeze_encoder (trainer.py:253-287): parameters that had no gradient when the buckets were cut
    take the late path for one step, the buckets are re-learnt, and the ranks never diverge."""
    from _dp_gloo_worker import Toy, data
    res = _run(tmp_path, "unfreeze")
    a, b = res["params"]
    assert torch.equal(a, b)                                   # replicas identical after training through the unfreeze
    log = res["log"]
    assert all(l["a.weight"] is None for l in log[:3]) and all(l["a.weight"] is not None for l in log[3:])
    assert res["stats"]["late"] == 2 and res["stats"]["recut"] == 1
    # single-process reference: same schedule on the global batch
    torch.manual_seed(5)
    ref = Toy()
    for p in ref.a.parameters():
        p.requires_grad = False
    opt = torch.optim.SGD(ref.parameters(), lr=0.05)
    X, Y = data()
    for step in range(6):
        if step == 3:
            for p in ref.a.parameters():
                p.requires_grad_(True)
        opt.zero_grad(set_to_none=True)
        ((ref(X) - Y) ** 2).mean().backward()
        opt.step()
        for k, p in ref.named_parameters():
            if p.grad is not None:
                assert torch.allclose(log[step][k], p.grad, rtol=1e-4, atol=1e-6), (step, k)
    assert torch.allclose(a, torch.cat([p.detach().reshape(-1) for p in ref.parameters()]), rtol=1e-4, atol=1e-6)


def test_sharded_adamw_equals_replicated_adamw(tmp_path):
    """reduce-scatter -> AdamW on the rank's shard -> all-gather equals torch.optim.AdamW on the global batch
    (decay / no-decay groups, an lr change between steps, a parameter without gradient), state_dict in torch's layout."""
    from _dp_gloo_worker import Toy, data
    res = _run(tmp_path, "sharded")
    assert res["raised"]                                       # the real update is CUDA-only: CPU tensors are refused
    assert torch.equal(res["ranks"][0], res["ranks"][1])
    assert len(res["shards"]) >= 2 and all(n >= 1 for _, n in res["shards"])
    ref = Toy()
    ref.load_state_dict(res["init"])
    decay = [p for k, p in ref.named_parameters() if p.dim() > 1 and not k.startswith("dead.")]
    no_decay = [p for k, p in ref.named_parameters() if p.dim() <= 1 and not k.startswith("dead.")]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.1}, {"params": no_decay, "weight_decay": 0.0}],
                            lr=1e-2, betas=(0.9, 0.95), eps=1e-8)
    X, Y = data()
    for step in range(4):
        if step == 2:
            for g in opt.param_groups:
                g["lr"] = 5e-3
        opt.zero_grad()
        ((ref(X) - Y) ** 2).mean().backward()
        opt.step()
    for k, v in ref.state_dict().items():
        assert torch.allclose(res["final"][k], v, rtol=2e-5, atol=2e-6), k
    assert torch.equal(res["final"]["dead.weight"], res["init"]["dead.weight"])
    want = opt.state_dict()
    got = res["opt_state"]
    assert [g["params"] for g in got["param_groups"]] == [g["params"] for g in want["param_groups"]]
    for i, st in want["state"].items():
        assert torch.allclose(got["state"][i]["exp_avg"], st["exp_avg"], rtol=2e-5, atol=1e-7)
        assert torch.allclose(got["state"][i]["exp_avg_sq"], st["exp_avg_sq"], rtol=2e-5, atol=1e-9)
        assert float(got["state"][i]["step"]) == float(st["step"]) == 4.0
