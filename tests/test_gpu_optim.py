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


def _small_model(seed=0):
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys
    torch.manual_seed(seed)
    m = MSUNetSys(img_size=64, embed_dim=32, depths=[2, 2, 2, 2], num_heads=[1, 2, 4, 8], drop_path_rate=0.0).to(DEV)
    m.set_precision("bf16")
    return m.train()


def _batch(seed=3, B=2, S=64):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, S, S, generator=g).to(DEV)
    y = (torch.rand(B, S, S, generator=g) > 0.7).float().to(DEV)
    return x, y


def test_weight_shadows_follow_the_optimizer():
    """Training in bf16 with FusedAdamW tracks the same run with torch.optim.AdamW: the raw-pointer update bumps the version
    counters, so the bf16 weight shadows are re-derived (one msu_refresh_shadows launch) before the next forward."""
    import semantic_segmentation_of_stylegan2_artifacts_b200 as pkg
    from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss
    from semantic_segmentation_of_stylegan2_artifacts_b200.optim import FusedAdamW
    ma, mb, mc = _small_model(0), _small_model(0), _small_model(0)
    oa = FusedAdamW(ma.parameters(), lr=2e-3, weight_decay=0.01)
    ob = torch.optim.AdamW(mb.parameters(), lr=2e-3, weight_decay=0.01, foreach=False, fused=False)
    oc = torch.optim.AdamW(mc.parameters(), lr=2e-3, weight_decay=0.01, fused=True)   # never bumps the version counters
    crit = DynamicLoss()
    x, y = _batch()
    la, lb, lc = [], [], []
    for _ in range(4):
        for m, o, ls in ((ma, oa, la), (mb, ob, lb), (mc, oc, lc)):
            o.zero_grad(set_to_none=True)
            loss = crit(m(x), y)
            loss.backward()
            o.step()
            ls.append(float(loss.detach()))
    assert abs(la[0] - lb[0]) < 1e-6                       # identical first step
    assert la[3] < la[0] - 1e-3                            # the updates reach the forward pass (stale shadows would freeze the loss)
    assert all(abs(a - b) < 0.02 * abs(b) + 1e-3 for a, b in zip(la, lb)), (la, lb)
    assert all(abs(c - b) < 0.02 * abs(b) + 1e-3 for c, b in zip(lc, lb)), (lc, lb)
    mb.eval(); mc.eval()
    with torch.no_grad():                                  # the first forward after training sees the last update too
        assert (mb(x).float() - mc(x).float()).abs().max() < 0.05
    # the refresh is one launch: a forward right after an optimizer step launches no per-tensor msu_prep_weight kernels
    n0 = pkg.launch_count()
    with torch.no_grad():
        ma(x)
    n_after_step = pkg.launch_count() - n0
    n0 = pkg.launch_count()
    with torch.no_grad():
        ma(x)
    assert n_after_step == (pkg.launch_count() - n0) + 1


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_refresh_shadows_bit_exact_vs_per_tensor_prep_and_under_graph_replay(prec):
    from semantic_segmentation_of_stylegan2_artifacts_b200 import functional as Fn, ops
    from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss
    m = _small_model(1)
    m.set_precision(prec)
    crit = DynamicLoss()
    x, y = _batch(5)
    for _ in range(2):                                     # registers forward (modes 0, 2, 5) and backward (modes 1, 3) shadows
        m.zero_grad(set_to_none=True)
        crit(m(x), y).backward()
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.25).add_(0.01)                        # in-place: version counters move, storage stays
        logits = m(x)                                      # one refresh launch covers every registered shadow
    seen = set()
    for (pid, mode, dt), ent in list(Fn._shadows.items()):
        p = ent[0]()
        if p is None or not any(p is q for q in m.parameters()):
            continue
        R, Cc = ent[3]
        want = ops.prep_weight(mode, p, R, Cc, tuple(ent[2].shape), dt)
        assert ent[1] == (p._version, p.data_ptr()) and torch.equal(ent[2], want), (mode, tuple(p.shape))
        seen.add(mode)
    assert seen >= ({0, 1, 2, 3, 5} if prec == "bf16" else {2, 3, 5})      # fp32 mode reads the linear weights directly
    # graph capture: the refresh is recorded, so a replay after an in-place weight change equals the eager forward after it
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s), torch.no_grad():
        m(x)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            out = m(x)
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, logits)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(0.9)
    g.replay()
    torch.cuda.synchronize()
    with torch.no_grad():
        eager = m(x)
    assert torch.equal(out, eager) and not torch.equal(eager, logits)


def test_graphed_step_matches_eager_training():
    """graphs.GraphedStep (fwd + DynamicLoss + bwd replayed from one CUDA graph, optimizer eager) follows the eager loop: same
    losses step by step, gradients in p.grad, zero_grad(set_to_none=True) between steps is harmless, shape changes are refused."""
    from semantic_segmentation_of_stylegan2_artifacts_b200.graphs import GraphedStep
    from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss
    from semantic_segmentation_of_stylegan2_artifacts_b200.optim import FusedAdamW
    ma, mb = _small_model(2), _small_model(2)
    oa = FusedAdamW(ma.parameters(), lr=2e-3, weight_decay=0.01)
    ob = FusedAdamW(mb.parameters(), lr=2e-3, weight_decay=0.01)
    crit = DynamicLoss()
    batches = [_batch(10 + i) for i in range(4)]
    sd0 = {k: v.clone() for k, v in ma.state_dict().items()}
    step = GraphedStep(ma, crit, *batches[0], warmup=2)
    ma.load_state_dict(sd0)                                # the warm-up ran no optimizer, but be explicit about the start point
    la, lb = [], []
    for x, y in batches:
        la.append(float(step(x, y)))
        has_a = [p.grad is not None for p in ma.parameters()]
        oa.step()
        oa.zero_grad(set_to_none=True)
        ob.zero_grad(set_to_none=True)
        loss = crit(mb(x), y)
        loss.backward()
        assert has_a == [p.grad is not None for p in mb.parameters()] and sum(has_a) > 100   # (dead decoder branches have none)
        ob.step()
        lb.append(float(loss.detach()))
    assert abs(la[0] - lb[0]) < 1e-6 and la[-1] != la[0]
    assert all(abs(a - b) < 0.02 * abs(b) + 1e-3 for a, b in zip(la, lb)), (la, lb)
    with pytest.raises(ValueError):
        step(batches[0][0][:1], batches[0][1][:1])
    with pytest.raises(RuntimeError):
        GraphedStep(ma, crit, batches[0][0].cpu(), batches[0][1].cpu())


def test_fused_adamw_under_grad_scaler_matches_torch():
    """trainer.py:182, 314-316: `scaler.scale(loss).backward(); scaler.step(optimizer); scaler.update()`.  FusedAdamW takes the
    scale through the GradScaler protocol (`grad_scale` / `found_inf`: 1/scale is applied inside the kernel, an overflow step is
    skipped together with its step count) and must track stock torch.optim.AdamW driven by its own scaler, overflow included."""
    from semantic_segmentation_of_stylegan2_artifacts_b200.optim import FusedAdamW
    pa, _ = _params(3)
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = FusedAdamW(pa, lr=2e-3, weight_decay=0.01)
    ob = torch.optim.AdamW(pb, lr=2e-3, weight_decay=0.01, foreach=False, fused=False)
    sa, sb = torch.amp.GradScaler("cuda", init_scale=1024.0), torch.amp.GradScaler("cuda", init_scale=1024.0)
    for step in range(5):
        torch.manual_seed(7 + step)
        ws = [torch.randn_like(p) for p in pa]
        for ps, opt, sc in ((pa, oa, sa), (pb, ob, sb)):
            opt.zero_grad(set_to_none=True)
            loss = sum((p * w).sum() + (p * p).sum() for p, w in zip(ps, ws))
            sc.scale(loss).backward()
            if step == 2:
                ps[0].grad[0, 0] = float("inf")          # overflow: both must skip this step and halve the scale
            sc.step(opt)
            sc.update()
    torch.cuda.synchronize()
    assert sa.get_scale() == sb.get_scale() == 512.0
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=3e-6, atol=1e-7), float((a - b).abs().max())
    assert all(float(s["step"]) == 4.0 for s in oa.state_dict()["state"].values())
