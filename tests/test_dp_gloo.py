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


def test_two_rank_gradients_equal_global_batch(tmp_path):
    from _dp_gloo_worker import Toy, data
    out = str(tmp_path / "rank0.pt")
    port = str(_free_port())
    worker = os.path.join(ROOT, "tests", "_dp_gloo_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", port, out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=180) == 0
    res = torch.load(out, weights_only=False)
    grads, sd, buckets, c, s = res["grads"], res["sd"], res["buckets"], res["c"], res["s"]
    ref = Toy()
    ref.load_state_dict(sd)
    assert all(not k.startswith("module.") for k in sd)   # checkpoints keep the reference's key names
    X, Y = data()
    ((ref(X) - Y) ** 2).mean().backward()
    assert len(buckets) >= 2                      # the tiny bucket size forces several buckets
    for step_grads in grads:
        for k, p in ref.named_parameters():
            if k.startswith("dead."):
                assert step_grads[k] is None and p.grad is None
            else:
                assert torch.allclose(step_grads[k], p.grad, rtol=1e-5, atol=1e-7), k
    assert c.shape == (4, 4) and c[:2].eq(0).all() and c[2:].eq(1).all()
    assert s.shape == (4, 8) and s[2:].eq(1.0).all()
