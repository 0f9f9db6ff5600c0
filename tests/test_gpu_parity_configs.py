"""Parity of the CUDA path AT THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs[1..3]: T96, 512x512, bf16, drop_path 0.1;
configs[3]: 1024x1024 inference + counting), with stochastic depth composed into the model.

The stochastic-depth noise is the one the REFERENCE drew (recorded by oracle/make_golden.py:run_sd_case through forward hooks
on torchvision's StochasticDepth, TV:ops/stochastic_depth.py:35-44) and is injected into both the CPU oracle
(`sd_noise=`) and the CUDA model (`MSUNetSys.inject_drop_path_noise`).  References, in order of authority:
  * tests/golden/t32_160_sd.npz / t96_512_sd.npz — outputs of the reference itself (network/model_parts.py:850-855);
  * the CPU oracle run live on the same inputs (full logits + every gradient tensor).
Tolerances (BASELINE.json north_star): fp32 mode max-rel <= 1e-3 on logits and gradients, loss rel <= 1e-4;
bf16 mode logits <= 3e-2 max-rel / 2e-2 rel-L2, gradient tensors rel-L2 <= 5e-2 (8e-2 bias tables), loss <= 2e-3; counts bit exact.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import msunet_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def relmax(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rell2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def load_sd_noise(g):
    return {str(k): (torch.from_numpy(n[0].copy()), torch.from_numpy(n[1].copy())) for k, n in zip(g["noise_names"], g["noise"])}


def build(cfg, prec, drop_path):
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys
    m = MSUNetSys(img_size=cfg.img_size, embed_dim=cfg.embed_dim, depths=list(cfg.depths), num_heads=list(cfg.num_heads),
                  drop_path_rate=drop_path)
    m.load_state_dict(O.make_weights(cfg), strict=True)
    return m.set_precision(prec).to(DEV)


SD_CASES = {"t32_160_sd": (O.T32, 160, 4), "t96_512_sd": (O.T96, 512, 2)}
_oracle_cache = {}


def oracle_step(name):
    """One CPU oracle training step per fixture (shared by the fp32 and bf16 runs)."""
    if name not in _oracle_cache:
        kw, img, batch = SD_CASES[name]
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        cfg = O.Cfg(img_size=img, **kw)
        x, y = O.make_inputs(cfg, batch)
        torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
        _oracle_cache[name] = (cfg, g, x, y) + tuple(O.train_step(O.make_weights(cfg), x, y, cfg, sd_noise=load_sd_noise(g)))
    return _oracle_cache[name]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(SD_CASES))
def test_training_step_with_stochastic_depth(name, prec):
    """t96_512_sd is the benchmarked shape: stage 0/1 take the fused_sd branch (HW >= 4096: row scale in the dgrad epilogue and
    the split-K reduce), stages 2/3 the gathered-rows branch; padded maps 133/70/35/21; conv tiles at W = 512; 5-D TMA stores
    at H = 128.  t32_160_sd drops 16 of 128 branches on small padded / shifted / shift-disabled maps."""
    from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss
    cfg, g, x, y, ref_logits, ref_loss, ref_grads = oracle_step(name)
    f32 = prec == "fp32"
    m = build(cfg, prec, float(g["drop_path_rate"])).train()
    m.inject_drop_path_noise(load_sd_noise(g))
    logits = m(x.to(DEV))
    loss = DynamicLoss(alpha=0.2, beta=0.8, tversky_bce_mix=0.45)(logits, y.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    # --- against the reference's own outputs
    st = int(g["stride"])
    gl = torch.from_numpy(g["logits_strided"])
    assert relmax(logits.float()[:, :, ::st, ::st], gl) < (1e-3 if f32 else 3e-2)
    assert abs(loss.item() - float(g["loss"])) < (1e-4 if f32 else 2e-3) * float(g["loss"])
    params = dict(m.named_parameters())
    for k, n in zip(g["grad_names"], g["grad_norms"]):
        gn = params[str(k)].grad.double().norm().item()
        tol = (1e-3 if f32 else 5e-2) * (3 if "attn.qkv.bias" in str(k) else 1)
        assert abs(gn - n) < tol * n + 1e-7, (k, gn, n)
    # --- against the oracle: full logits and EVERY gradient tensor element-wise
    assert relmax(logits.float(), ref_logits) < (1e-3 if f32 else 3e-2)
    assert rell2(logits.float(), ref_logits) < (1e-4 if f32 else 2e-2)
    assert abs(loss.item() - ref_loss.item()) < (1e-4 if f32 else 2e-3) * abs(ref_loss.item())
    worst = ("", 0.0)
    for k, r in ref_grads.items():
        if k.startswith(O.DEAD_PREFIXES):
            assert params[k].grad is None, k
            continue
        got = params[k].grad
        assert got is not None, k
        if "attn.qkv.bias" in k:        # the K third is mathematically zero: compare on the tensor's scale
            e = float((got.cpu() - r).abs().max() / r.abs().max())
            assert e < (1e-3 if f32 else 5e-2), (k, e)
            continue
        e = relmax(got, r) if f32 else rell2(got, r)
        tol = 1e-3 if f32 else (8e-2 if "relative_position_bias_table" in k else 5e-2)
        if e > worst[1]:
            worst = (k, e)
        assert e < tol, (k, e)
    # --- the noise really was applied: a second forward with other noise differs
    other = O.draw_sd_noise(cfg, x.shape[0], float(g["drop_path_rate"]), seed=5)
    m.inject_drop_path_noise(other)
    with torch.no_grad():
        l2 = m(x.to(DEV))
    assert relmax(l2.float(), ref_logits) > (1e-3 if f32 else 3e-2)
    print(f"{name}[{prec}] logits relmax {relmax(logits.float(), ref_logits):.2e} rel-L2 {rell2(logits.float(), ref_logits):.2e} "
          f"loss rel {abs(loss.item() - ref_loss.item()) / abs(ref_loss.item()):.2e} worst grad {worst[0]} {worst[1]:.2e}")


def test_stochastic_depth_draws_follow_the_block_probabilities():
    """Without injection the model draws Bernoulli(1-p)/(1-p) per sample and block (row mode): value set and drop rate."""
    cfg = O.Cfg(img_size=64, **O.T32)
    m = build(cfg, "bf16", 0.5).train()
    probs = O.block_drop_probs(cfg, 0.5)
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import SwinTransformerBlock
    torch.manual_seed(1)
    m._draw_drop_path(512, torch.device(DEV))
    zeros, total = 0, 0
    for name, b in m.named_modules():
        if isinstance(b, SwinTransformerBlock):
            assert abs(b.sd_prob - probs[name]) < 1e-7
            if b.sd_prob > 0:
                for v in b._sd_pool:
                    vv = v.cpu()
                    assert bool(((vv == 0) | ((vv - 1.0 / (1.0 - b.sd_prob)).abs() < 1e-5)).all()), (name, vv.unique())
                    frac = float((v == 0).float().mean())
                    assert abs(frac - b.sd_prob) < 0.08, (name, frac, b.sd_prob)
                    zeros += int((v == 0).sum()); total += v.numel()
    assert total > 0 and 0.15 < zeros / total < 0.4


@pytest.mark.parametrize("kw,tag", [(O.T32, "T32"), (O.T96, "T96")])
def test_inference_and_counts_at_1024(kw, tag):
    """BASELINE.json configs[3]: 1024x1024 eval forward (padded maps 259/133/70/35) + sigmoid / threshold(0.5) / TP-FP-FN-TN.
    Counts must be bit-exact against the reference formulas (scripts/validation_functions.py:106-108, 214-309 restated in the
    oracle) applied to the SAME logits; the logits themselves are checked against the CPU oracle forward (fp32 mode for T32,
    bf16 for both)."""
    from semantic_segmentation_of_stylegan2_artifacts_b200 import ops
    cfg = O.Cfg(img_size=1024, **kw)
    x, y = O.make_inputs(cfg, 1, real_last=False)
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    with torch.no_grad():
        ref = O.forward(O.make_weights(cfg), x, cfg)
    for prec in (("fp32", "bf16") if tag == "T32" else ("bf16",)):
        m = build(cfg, prec, 0.1).eval()
        with torch.no_grad():
            logits = m(x.to(DEV))
        f32 = prec == "fp32"
        assert relmax(logits.float(), ref) < (1e-3 if f32 else 3e-2), prec
        assert rell2(logits.float(), ref) < (1e-4 if f32 else 2e-2), prec
        lab = (y * 255).to(DEV)
        counts, soft, pred = ops.metrics(logits.view(1, -1), lab.view(1, -1), None, True, 0.5, want_pred=True)
        torch.cuda.synchronize()
        # the reference thresholds sigmoid(logits) in the logits' dtype: restate on the host from the same logits
        p_host = torch.sigmoid(logits)[0, 0].cpu()          # torch's own expression on the same logits / dtype
        pb = (p_host > 0.5).numpy()
        gt = (lab.cpu()[0] > 0).numpy()
        assert tuple(int(v) for v in counts[0].tolist()) == O.confusion_counts(pb, gt), prec
        want = O.soft_sums(p_host.float().numpy(), gt)
        np.testing.assert_allclose(soft[0].cpu().numpy(), np.array(want), rtol=1e-5)
        # the validation loop's mode: fp32 probabilities from the same logits
        c32, s32, p32 = ops.metrics(logits.view(1, -1), lab.view(1, -1), None, True, 0.5, want_pred=True, prob_f32=True)
        p_host32 = torch.sigmoid(logits.float())[0, 0].cpu()
        assert torch.equal(p32.view(1024, 1024).cpu(), p_host32)
        assert tuple(int(v) for v in c32[0].tolist()) == O.confusion_counts((p_host32 > 0.5).numpy(), gt), prec


def test_per_sample_loss_decides_the_label_scale_per_image():
    """DynamicLoss.per_sample (batched validation): image b is a batch of one for the reference (validation_functions.py:89-104),
    so a {0,1}-labelled image next to a {0,255}-labelled one keeps its own interpretation (loss/DynamicLoss.py:87-88)."""
    from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss
    g = torch.Generator().manual_seed(3)
    lg = torch.randn(4, 1, 64, 64, generator=g) * 2
    t = (torch.rand(4, 64, 64, generator=g) > 0.8).float()
    t[1] *= 255.0
    t[2] = 0
    t[3] *= 255.0
    crit = DynamicLoss(alpha=0.2, beta=0.8, tversky_bce_mix=0.45)
    for dt in (torch.float32, torch.bfloat16):
        got = crit.per_sample(lg.to(DEV).to(dt), t.to(DEV)).cpu()
        want = torch.stack([O.dynamic_loss(lg[i:i + 1].to(dt).float(), t[i:i + 1], 0.2, 0.8, 0.45) for i in range(4)])
        assert float((got - want).abs().max()) < 2e-6 * float(want.abs().max()) + 1e-6, (dt, got, want)
        # forward() keeps the reference's batch-global decision
        whole = crit(lg.to(DEV).to(dt), t.to(DEV)).item()
        assert abs(whole - O.dynamic_loss(lg.to(dt).float(), t, 0.2, 0.8, 0.45).item()) < 1e-5
