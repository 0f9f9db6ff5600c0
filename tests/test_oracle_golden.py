"""The oracle restatement vs fixtures produced by RUNNING THE REFERENCE (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import msunet_oracle as O

CASES = {"t32_160": (O.T32, 160, 2), "t96_224": (O.T96, 224, 2)}


def relmax(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("name", list(CASES))
def test_forward_backward_matches_reference(name):
    kw, img, batch = CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = O.Cfg(img_size=img, **kw)
    sd = O.make_weights(cfg)
    assert list(sd.keys()) == list(g["sd_keys"])
    assert [",".join(map(str, v.shape)) for v in sd.values()] == list(g["sd_shapes"])
    x, y = O.make_inputs(cfg, batch)
    logits, loss, grads = O.train_step(sd, x, y, cfg)
    assert relmax(logits.numpy(), g["logits"]) < 2e-5
    assert abs(loss.item() - float(g["loss"])) < 1e-6 * abs(float(g["loss"])) + 1e-7
    assert abs(O.dynamic_loss(logits, y * 255, 0.2, 0.8, 0.45).item() - float(g["loss_255"])) < 1e-6
    dead = set(g["dead"])
    assert dead == {k for k in grads if k.startswith(O.DEAD_PREFIXES)}
    for k in dead:
        assert grads[k] is None
    for k, n, s in zip(g["grad_names"], g["grad_norms"], g["grad_sums"]):
        gn = grads[k].double().norm().item()
        if "attn.qkv.bias" in k:  # K-third is mathematically zero: rounding noise
            assert abs(gn - n) < 1e-4 * n + 1e-5, k
        else:
            assert abs(gn - n) < 2e-4 * n + 1e-7, (k, gn, n)
    for key in g.files:
        if key.startswith("grad::"):
            k = key[6:]
            tol = 5e-4 if "qkv.bias" not in k else 2e-3
            assert relmax(grads[k].numpy(), g[key]) < tol, k


SD_CASES = {"t32_160_sd": (O.T32, 160, 4), "t96_512_sd": (O.T96, 512, 2)}


def load_sd_noise(g):
    """{block name: (noise1 [B], noise2 [B])} as the reference drew it (recorded by oracle/make_golden.py:run_sd_case)."""
    return {str(k): (torch.from_numpy(n[0].copy()), torch.from_numpy(n[1].copy())) for k, n in zip(g["noise_names"], g["noise"])}


@pytest.mark.parametrize("name", list(SD_CASES))
def test_stochastic_depth_matches_reference(name):
    """Training forward + backward WITH stochastic depth (drop_path 0.3 on small padded maps; drop_path 0.1 at the benchmarked
    512x512 T96 shape): the oracle fed the noise the reference drew reproduces the reference's logits, loss and gradients."""
    kw, img, batch = SD_CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = O.Cfg(img_size=img, **kw)
    sd = O.make_weights(cfg)
    noise = load_sd_noise(g)
    probs = O.block_drop_probs(cfg, float(g["drop_path_rate"]))
    assert set(noise) == {k for k, p in probs.items() if p > 0}
    for k, (n1, n2) in noise.items():       # every value is 0 or 1/(1-p) of ITS block
        for v in torch.cat([n1, n2]).tolist():
            assert v == 0.0 or abs(v - 1.0 / (1.0 - probs[k])) < 1e-6, (k, v)
    x, y = O.make_inputs(cfg, batch)
    logits, loss, grads = O.train_step(sd, x, y, cfg, sd_noise=noise)
    st = int(g["stride"])
    assert relmax(logits[:, :, ::st, ::st].numpy(), g["logits_strided"]) < 2e-5
    assert abs(logits.double().norm().item() - float(g["logits_l2"])) < 1e-5 * float(g["logits_l2"])
    assert abs(logits.double().sum().item() - float(g["logits_sum"])) < 1e-5 * float(g["logits_abs_sum"])
    assert abs(loss.item() - float(g["loss"])) < 1e-6 * abs(float(g["loss"])) + 1e-7
    for k, n in zip(g["grad_names"], g["grad_norms"]):
        gn = grads[k].double().norm().item()
        assert abs(gn - n) < (1e-4 if "attn.qkv.bias" in k else 2e-4) * n + 1e-5, (k, gn, n)
    for key in g.files:
        if key.startswith("grad::"):
            k = key[6:]
            assert relmax(grads[k].numpy(), g[key]) < (5e-4 if "qkv.bias" not in k else 2e-3), k
    # and the noise matters: without it the logits differ
    assert relmax(O.forward(sd, x[:1], cfg)[:, :, ::st, ::st].numpy(), g["logits_strided"][:1]) > 1e-3


def test_draw_sd_noise_is_row_mode_bernoulli():
    cfg = O.Cfg(img_size=64, **O.T32)
    n = O.draw_sd_noise(cfg, 64, 0.5, seed=3)
    probs = O.block_drop_probs(cfg, 0.5)
    assert "layers.0.blocks.0" not in n and probs["layers.0.blocks.0"] == 0.0          # dpr starts at 0
    assert abs(probs["layers.3.blocks.1"] - 0.5) < 1e-7 and probs["layers_up.3.blocks.1"] == probs["layers.0.blocks.1"]
    for k, (a, b) in n.items():
        keep = 1.0 - probs[k]
        for v in (a, b):
            assert set(v.tolist()) <= {0.0, float(torch.tensor(1.0) / keep)}
    allv = torch.cat([torch.cat(v) for v in n.values()])
    assert 0.1 < float((allv == 0).float().mean()) < 0.45


def test_dead_branches_do_not_change_logits():
    cfg = O.Cfg(img_size=96, **O.T32)
    sd = O.make_weights(cfg)
    x, _ = O.make_inputs(cfg, 1)
    a = O.forward(sd, x, cfg, run_dead=False)
    b = O.forward(sd, x, cfg, run_dead=True)
    assert torch.equal(a, b)


def test_window_attention_matches_torchvision():
    from torchvision.models.swin_transformer import SwinTransformerBlock
    for (H, W, shift, C, nH) in [(14, 14, 3, 32, 1), (16, 16, 3, 64, 2), (16, 16, 0, 64, 2),
                                 (7, 7, 3, 32, 1), (10, 12, 3, 32, 1), (5, 5, 3, 64, 2)]:
        torch.manual_seed(H * 100 + W + shift)
        blk = SwinTransformerBlock(C, nH, [7, 7], [shift, shift]).eval()
        with torch.no_grad():
            for p in blk.parameters():
                p.add_(0.05 * torch.randn_like(p))
        sd = {"b." + k: v for k, v in blk.state_dict().items()}
        x = torch.randn(2, H, W, C)
        with torch.no_grad():
            ref = blk(x)
            got = O.swin_block(x, sd, "b", nH, shift)
        assert relmax(got.numpy(), ref.numpy()) < 1e-5, (H, W, shift)


def test_loss_known_answers():
    g = np.load(os.path.join(GOLDEN, "loss_cases.npz"))
    for tag in ("a", "b"):
        lg = torch.from_numpy(g[tag + "_logits"]).requires_grad_(True)
        t = torch.from_numpy(g[tag + "_target"])
        for (al, be, mx) in ((0.4, 0.6, 0.5), (0.2, 0.8, 0.45)):
            lg.grad = None
            l = O.dynamic_loss(lg, t, al, be, mx)
            l.backward()
            key = f"{tag}_{al}_{be}_{mx}"
            assert abs(l.item() - float(g[key + "_loss"])) < 2e-7
            assert relmax(lg.grad.numpy(), g[key + "_grad"]) < 1e-5


@pytest.mark.parametrize("name", list(CASES))
def test_metrics_match_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    kw, img, batch = CASES[name]
    cfg = O.Cfg(img_size=img, **kw)
    _, y = O.make_inputs(cfg, batch)
    for i in range(batch):
        pred = torch.sigmoid(torch.from_numpy(g["logits"][i, 0])).numpy()
        pb = pred > 0.5
        gt = y[i].numpy() > 0
        if gt.any():
            r = O.metrics_fake(pb, pred, gt)
            assert np.array_equal(np.array(r[6]), g[f"fake{i}_cm_bin"])  # bit-exact integers
            np.testing.assert_allclose(list(r[:6]) + [r[8], r[9]], g[f"fake{i}_scalars"], rtol=2e-6)
            np.testing.assert_allclose(np.array(r[7]), g[f"fake{i}_cm_soft"], rtol=2e-6)
        else:
            cb, cs, acc, fpr = O.metrics_real(pb, pred, gt)
            assert np.array_equal(np.array(cb), g[f"real{i}_cm_bin"])
            np.testing.assert_allclose(np.array(cs), g[f"real{i}_cm_soft"], rtol=2e-6)
            np.testing.assert_allclose([acc, fpr], g[f"real{i}_scalars"], rtol=1e-12)


def test_staging_matches_reference_transform():
    """oracle.stage_batch vs tensors returned by the reference's RandomGenerator (oracle/make_staging_golden.py): bit-exact."""
    g = np.load(os.path.join(GOLDEN, "staging.npz"))
    img, lab = O.stage_batch(g["images"], g["labels"], g["flips"])
    assert img.dtype == np.float32 and lab.dtype == np.float32
    assert np.array_equal(img, g["out_image"]) and np.array_equal(lab, g["out_label"])
    assert g["flips"].any() and not g["flips"].all()


def test_attention_dropout_mask_restatement():
    """oracle.attn_drop_keep (the counter-hash mask the CUDA kernels regenerate in forward and backward): integer hash vs plain
    Python arithmetic, keep rate, determinism, seed sensitivity, p = 0."""
    def lowbias32(x):
        x ^= x >> 16; x = (x * 0x7FEB352D) & 0xFFFFFFFF
        x ^= x >> 15; x = (x * 0x846CA68B) & 0xFFFFFFFF
        return x ^ (x >> 16)
    xs = [0, 1, 2, 0xFFFFFFFF, 0x12345678, 0x9E3779B1]
    assert [int(v) for v in O._lowbias32(np.array(xs, dtype=np.uint64))] == [lowbias32(x) for x in xs]
    p, s0, s1 = 0.05, 123456789, 987654321
    k = O.attn_drop_keep(6, 3, p, s0, s1)
    assert k.dtype == torch.bool and tuple(k.shape) == (6, 3, 49, 49)
    thr = int(p * 65536.0 + 0.5)
    for (w, h, i, j) in ((0, 0, 0, 0), (5, 2, 48, 48), (3, 1, 17, 30), (2, 0, 9, 11)):      # element by element, from the definition
        rowkey = (w * 3 + h) * 49 + i
        x = ((((rowkey << 5) & 0xFFFFFFFF) + (j >> 1) + s0) & 0xFFFFFFFF) * 0x9E3779B1 & 0xFFFFFFFF
        hsh = lowbias32(x ^ s1)
        half = (hsh >> 16) if (j & 1) else (hsh & 0xFFFF)
        assert bool(k[w, h, i, j]) == (half >= thr)
    assert abs(float((~k).float().mean()) - p) < 0.01
    assert torch.equal(k, O.attn_drop_keep(6, 3, p, s0, s1)) and not torch.equal(k, O.attn_drop_keep(6, 3, p, s0 + 1, s1))
    assert bool(O.attn_drop_keep(2, 1, 0.0, 1, 2).all())
