"""Fused multi-tensor AdamW against torch.optim.AdamW on identical parameters and gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(96, 96), (288,), (384, 96), (96,), (3, 5, 7), (1,), (8192 * 2 + 3,), (169, 3)]
    ps = [torch.randn(*s, generator=g).to(DEV).requires_grad_(True) for s in shapes]
    # parameters that are odd-offset views of a flat buffer (the DP wrapper points .grad at bucket slices): scalar path
    flat = torch.randn(1000, generator=g).to(DEV)
    return ps, flat


def test_fused_adamw_matches_torch():
    from semantic_segmentation_of_stylegan2_artifacts_b200.optim import FusedAdamW
    pa, _ = _params(1)
    pb = [p.detach().clone().requires_grad_(True) for p in pa]

    def groups(ps):
        return [{"params": [p for p in ps if p.ndim > 1], "weight_decay": 0.05},
                {"params": [p for p in ps if p.ndim <= 1], "weight_decay": 0.0}]

    oa = FusedAdamW(groups(pa), lr=3e-3, betas=(0.9, 0.98), eps=1e-8)
    ob = torch.optim.AdamW(groups(pb), lr=3e-3, betas=(0.9, 0.98), eps=1e-8, foreach=False, fused=False)
    gbuf = torch.zeros(4099, device=DEV)
    for step in range(6):
        torch.manual_seed(100 + step)
        for i, (a, b) in enumerate(zip(pa, pb)):
            gr = torch.randn_like(a) * (0.1 + step)
            if i == 3 and step == 2:
                a.grad = b.grad = None          # a parameter without gradient is skipped (its step counter does not advance)
                continue
            if i == 1:                          # gradient living at an odd offset of a flat bucket (4-byte aligned only)
                view = gbuf[3:3 + a.numel()].view_as(a)
                view.copy_(gr)
                a.grad = view
            else:
                a.grad = gr.clone()
            b.grad = gr.clone()
        for grp in oa.param_groups + ob.param_groups:
            grp["lr"] = 3e-3 * (1.0 - 0.1 * step)      # lr schedule (trainer.py:305-306)
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), float((a - b).abs().max())
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["state"].keys() == sb["state"].keys()
    for k in sa["state"]:
        assert set(sa["state"][k]) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"])
        assert torch.allclose(sa["state"][k]["exp_avg_sq"], sb["state"][k]["exp_avg_sq"], rtol=2e-6, atol=1e-12)
    # state interchange: a torch AdamW checkpoint loads into the fused optimizer and vice versa
    oa.load_state_dict(sb)
    ob.load_state_dict(sa)


def test_fused_adamw_rejects_cpu_and_amsgrad():
    from semantic_segmentation_of_stylegan2_artifacts_b200.optim import FusedAdamW
    with pytest.raises(NotImplementedError):
        FusedAdamW([torch.zeros(3, requires_grad=True)], amsgrad=True)
    p = torch.zeros(3, requires_grad=True)
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError):
        FusedAdamW([p]).step()
