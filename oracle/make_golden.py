"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (build container only).

Imports network/model_parts.py:MSUNetSys, loss/DynamicLoss.py:DynamicLoss and
scripts/validation_functions.py from /root/reference (with the timm / medpy stand-ins in
oracle/_shims), loads the deterministic weights of oracle.msunet_oracle.make_weights with
strict=True (which also pins the state_dict key/shape contract), runs forward + loss +
backward on CPU fp32 and stores the results.  The fixtures travel; the reference does not.

    python oracle/make_golden.py            # rewrites tests/golden/
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_shims"))
sys.path.insert(0, "/root/reference")

from oracle import msunet_oracle as O  # noqa: E402
from network.model_parts import MSUNetSys  # noqa: E402  (reference)
from loss.DynamicLoss import DynamicLoss  # noqa: E402  (reference)
from scripts import validation_functions as VF  # noqa: E402  (reference)

CASES = {
    # name: (cfg kwargs, img, batch)
    "t32_160": (O.T32, 160, 2),   # 40/20/10/5 -> padded 42/21/14/7: pad + shift + shift-disabled
    "t96_224": (O.T96, 224, 2),   # BASELINE.json configs[0]
}
FULL_GRADS = ("relative_position_bias_table", "qkv.bias", "norm1.weight", "output.weight",
              "up.norm.bias", "patch_embed.proj.weight", "concat_back_dim.3.bias")


def run_case(name, kw, img, batch):
    torch.manual_seed(0)
    cfg = O.Cfg(img_size=img, **kw)
    sd = O.make_weights(cfg)
    m = MSUNetSys(img_size=img, embed_dim=cfg.embed_dim, depths=list(cfg.depths),
                  num_heads=list(cfg.num_heads), window_size=7, drop_path_rate=0.0,
                  attn_drop_rate=0.0, drop_rate=0.0)
    m.load_state_dict(sd, strict=True)
    m.train()
    x, y = O.make_inputs(cfg, batch)
    crit = DynamicLoss(alpha=0.2, beta=0.8, tversky_bce_mix=0.45)
    logits = m(x)
    loss = crit(logits, y)
    loss.backward()
    out = {"logits": logits.detach().numpy(), "loss": np.float64(loss.item())}
    names, norms, sums, dead = [], [], [], []
    for k, p in m.named_parameters():
        if p.grad is None:
            dead.append(k)
            continue
        names.append(k)
        norms.append(p.grad.double().norm().item())
        sums.append(p.grad.double().sum().item())
        if any(s in k for s in FULL_GRADS) and p.grad.numel() <= 20000:
            out["grad::" + k] = p.grad.numpy()
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms)
    out["grad_sums"] = np.array(sums)
    out["dead"] = np.array(dead)
    out["sd_keys"] = np.array(list(m.state_dict().keys()))
    out["sd_shapes"] = np.array([",".join(map(str, v.shape)) for v in m.state_dict().values()])
    # 255-valued labels exercise the >127.5 branch (loss/DynamicLoss.py:87-88)
    out["loss_255"] = np.float64(crit(logits.detach(), y * 255).item())
    # metrics through the reference's own functions on these logits
    for i in range(batch):
        pred = torch.sigmoid(logits.detach()[i, 0])
        pb = pred > 0.5
        gt = y[i] > 0
        if gt.any():
            r = VF.calculate_metrics_fake(pb, pred, gt)
            out[f"fake{i}_scalars"] = np.array(list(r[:6]) + [r[8], r[9]], dtype=np.float64)
            out[f"fake{i}_cm_bin"] = np.array(r[6], dtype=np.int64)
            out[f"fake{i}_cm_soft"] = np.array(r[7], dtype=np.float64)
        else:
            cb, cs, acc, fpr = VF.calculate_metrics_real(pb, pred, gt)
            out[f"real{i}_cm_bin"] = np.array(cb, dtype=np.int64)
            out[f"real{i}_cm_soft"] = np.array(cs, dtype=np.float64)
            out[f"real{i}_scalars"] = np.array([acc, fpr], dtype=np.float64)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss", out["loss"], "logits sum", out["logits"].sum(), "dead", len(dead),
          "->", path, os.path.getsize(path) // 1024, "KiB")


SD_CASES = {
    # name: (cfg kwargs, img, batch, drop_path_rate, logits stride kept in the fixture)
    "t32_160_sd": (O.T32, 160, 4, 0.3, 1),    # stochastic depth on small padded / shifted maps (row gather path of the CUDA code)
    "t96_512_sd": (O.T96, 512, 2, 0.1, 8),    # BASELINE.json configs[1]/[2] shape and drop_path (config.yaml:23), 2 images
}


def run_sd_case(name, kw, img, batch, dpr, stride):
    """The reference in train mode WITH stochastic depth.  The Bernoulli noise torchvision draws inside
    `StochasticDepth.forward` (TV:ops/stochastic_depth.py:35-44) is recovered per call by forward hooks (a sample's
    branch is either all zero or scaled by 1/(1-p)) and stored, so the oracle / the CUDA path can be fed the same."""
    from torchvision.ops import StochasticDepth
    torch.manual_seed(20 + batch)
    cfg = O.Cfg(img_size=img, **kw)
    sd = O.make_weights(cfg)
    m = MSUNetSys(img_size=img, embed_dim=cfg.embed_dim, depths=list(cfg.depths), num_heads=list(cfg.num_heads),
                  window_size=7, drop_path_rate=dpr, attn_drop_rate=0.0, drop_rate=0.0)
    m.load_state_dict(sd, strict=True)
    m.train()
    seen = {}

    def hook(mod, inp, out, key=None):
        if mod.p == 0.0:
            return
        keep = 1.0 - mod.p
        o = out.detach().reshape(out.shape[0], -1)
        n = torch.where(o.abs().amax(1) > 0, torch.tensor(1.0 / keep), torch.tensor(0.0))
        assert torch.allclose(out, inp[0] * n.view(-1, 1, 1, 1), rtol=1e-6, atol=0)
        seen.setdefault(key, []).append(n.float())

    for mname, mod in m.named_modules():
        if isinstance(mod, StochasticDepth):
            assert mod.mode == "row"
            mod.register_forward_hook(lambda a, b, c, key=mname.rsplit(".", 1)[0]: hook(a, b, c, key))
    x, y = O.make_inputs(cfg, batch)
    crit = DynamicLoss(alpha=0.2, beta=0.8, tversky_bce_mix=0.45)
    logits = m(x)
    loss = crit(logits, y)
    loss.backward()
    probs = O.block_drop_probs(cfg, dpr)
    for k, v in seen.items():
        assert len(v) == 2 and abs(probs[k] - dict(m.named_modules())[k].stochastic_depth.p) < 1e-12, k
    assert set(seen) == {k for k, p in probs.items() if p > 0}
    lg = logits.detach()
    out = {"logits_strided": lg[:, :, ::stride, ::stride].numpy(), "stride": np.int64(stride),
           "logits_sum": np.float64(lg.double().sum().item()), "logits_abs_sum": np.float64(lg.double().abs().sum().item()),
           "logits_l2": np.float64(lg.double().norm().item()), "loss": np.float64(loss.item()),
           "drop_path_rate": np.float64(dpr), "noise_names": np.array(sorted(seen)),
           "noise": np.stack([torch.stack(seen[k]).numpy() for k in sorted(seen)])}          # [blocks, 2, B]
    names, norms, sums = [], [], []
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        names.append(k)
        norms.append(p.grad.double().norm().item())
        sums.append(p.grad.double().sum().item())
        if any(s in k for s in FULL_GRADS) and p.grad.numel() <= 20000:
            out["grad::" + k] = p.grad.numpy()
    out["grad_names"], out["grad_norms"], out["grad_sums"] = np.array(names), np.array(norms), np.array(sums)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    dropped = int((out["noise"] == 0).sum())
    print(name, "loss", out["loss"], "dropped branches", dropped, "of", out["noise"].size, "->", path,
          os.path.getsize(path) // 1024, "KiB")
    assert dropped >= 3, "pick another seed: the fixture should exercise dropped samples"


def loss_cases():
    """Loss-only known answers from the reference DynamicLoss on synthetic logits."""
    g = torch.Generator().manual_seed(7)
    out = {}
    for tag, (B, S) in {"a": (3, 33), "b": (4, 64)}.items():
        lg = (torch.randn(B, 1, S, S, generator=g) * 3).requires_grad_(True)
        t = (torch.rand(B, S, S, generator=g) > 0.8).float()
        t[0] = 0
        for (al, be, mx) in ((0.4, 0.6, 0.5), (0.2, 0.8, 0.45)):
            crit = DynamicLoss(alpha=al, beta=be, tversky_bce_mix=mx)
            lg.grad = None
            l = crit(lg, t)
            l.backward()
            key = f"{tag}_{al}_{be}_{mx}"
            out[key + "_loss"] = np.float64(l.item())
            out[key + "_grad"] = lg.grad.numpy().copy()
        out[tag + "_logits"] = lg.detach().numpy()
        out[tag + "_target"] = t.numpy()
    path = os.path.join(ROOT, "tests", "golden", "loss_cases.npz")
    np.savez_compressed(path, **out)
    print("loss cases ->", path)


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for n, (kw, img, b) in CASES.items():
        if not only or n in only:
            run_case(n, kw, img, b)
    for n, (kw, img, b, dpr, stride) in SD_CASES.items():
        if not only or n in only:
            run_sd_case(n, kw, img, b, dpr, stride)
    if not only or "loss_cases" in only:
        loss_cases()
