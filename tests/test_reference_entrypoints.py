"""The reference's OWN entry points — train.py:18-124 -> trainer.py:33-425 and test.py:29-165 — executed UNMODIFIED on the drop-in
overlay (BASELINE.json north_star: "train.py/trainer.py/test.py run unchanged").

`tools/install_reference.py` (run by `__graft_entry__.build()` where /root/reference exists) copies the reference's sources to
baseline/_ref/ (git-ignored, travels to the GPU box); `tests/ref_env/run_entry.py` runs a script from there with
sys.path = [dropin overlay, repo, tests/ref_env (stand-ins for yacs / tensorboardX / matplotlib / albumentations / timm, which this
image lacks, and for loss/SymmetricUnfiedFocalLoss_3.py, which the reference imports but does not ship), baseline/_ref].

Everything of the reference's loop runs as written: yacs config from a yaml, SegFace checkpoint loading, the name-based
decay / no-decay AdamW split (trainer.py:130-152), BatchPatternSampler (batch 2), `torch.amp.autocast(float16)` + `GradScaler`
(:182, 308-316), `zero_grad(set_to_none=True)`, per-epoch `calculate_metrics`, `best_model.pth` in the legacy serialisation
(:366-381, including the `.to('cpu')` / `.to(dev)` round trip), CosineLRScheduler, then test.py's `load_state_dict(strict=True)`,
CSV rows and heat-maps.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "baseline", "_ref")
S = 64

CFG = """BASE: []
DATA:
  BATCH_SIZE: 2
  DATA_PATH: "{data}"
  IMG_SIZE: {S}
  PIN_MEMORY: true
  NUM_WORKERS: 0
HARDWARE:
  N_GPU: 1
MODEL:
  PRETRAIN_WEIGHTS: "segface"
  PRETRAIN_SEGFACE: "{ckpt}"
  NUM_CLASSES: 1
  DROP_RATE: 0.0
  DROP_PATH_RATE: 0.1
  ATTN_DROP_RATE: 0.05
  FREEZE_ENCODER: false
  SWIN:
    PATCH_SIZE: 4
    IN_CHANS: 3
    EMBED_DIM: 32
    DEPTHS: [2, 2, 2, 2]
    DECODER_DEPTHS: [2, 2, 2, 2]
    NUM_HEADS: [1, 2, 4, 8]
    WINDOW_SIZE: 7
    MLP_RATIO: 4.0
TRAIN:
  MAX_EPOCHS: 3
  WARMUP_EPOCHS: 1
  WEIGHT_DECAY: 0.001
  BASE_LR: 0.002
  WARMUP_LR: 0.001
  MIN_LR: 0.0001
  TVERSKY_LOSS_ALPHA: 0.2
  TVERSKY_LOSS_BETA: 0.8
  LOSS_TVERSKY_BCE_MIX: 0.45
  SIG_THRESHOLD: 0.5
  EARLY_STOPPING_FLAG: false
OUTPUT_DIR: '{out}'
LIST_DIR: '{lists}'
SEED: 120
DETERMINISTIC: false
SHOW_PREDICTIONS: 2
SAVE_BEST_RUN: true
SAVE_LAST_RUN: true
DYNAMIC_LOADER: false
"""


def _make_dataset(root):
    """Layout of dataset/dataset.py:137-160: {fake,real}_images/<name>.png + {fake,real}_labels/<name>_mask.png; fake names start
    with "09" (trainer.py:463).  Fake images carry a bright rectangle under their mask, so the task is learnable in a few steps."""
    from PIL import Image
    rng = np.random.default_rng(0)
    for d in ("fake_images", "fake_labels", "real_images", "real_labels"):
        os.makedirs(os.path.join(root, "data", d), exist_ok=True)
    os.makedirs(os.path.join(root, "lists"), exist_ok=True)
    fake, real = [f"09{i:04d}" for i in range(10)], [f"00{i:04d}" for i in range(8)]
    for n in fake + real:
        img = (rng.random((S, S, 3)) * 80 + 40).astype(np.uint8)
        mask = np.zeros((S, S), np.uint8)
        kind = "fake" if n.startswith("09") else "real"
        if kind == "fake":
            y0, x0 = rng.integers(4, S - 28, 2)
            h, w = rng.integers(12, 24, 2)
            mask[y0:y0 + h, x0:x0 + w] = 255
            img[mask > 0] = np.clip(img[mask > 0].astype(np.int32) + 120, 0, 255).astype(np.uint8)
        Image.fromarray(img).save(os.path.join(root, "data", f"{kind}_images", n + ".png"))
        Image.fromarray(mask).save(os.path.join(root, "data", f"{kind}_labels", n + "_mask.png"))
    splits = {"fake_train": fake[:6], "real_train_all": real[:5], "real_train": real[:5], "train": fake[:6] + real[:5],
              "val": fake[6:8] + real[5:6], "test": fake[8:10] + real[6:8]}
    for k, v in splits.items():
        with open(os.path.join(root, "lists", k + ".txt"), "w") as f:
            f.write("\n".join(v) + "\n")
    return splits


def _run(script, report, args, cwd, env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    env.pop("PYTHONPATH", None)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_env", "run_entry.py"), script, report] + args,
                       cwd=cwd, env=env, capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0, p.stdout[-3000:] + "\n" + p.stderr[-6000:]
    return json.load(open(report))


@pytest.mark.skipif(not os.path.isdir(REF), reason="baseline/_ref is absent: run __graft_entry__.build() where /root/reference exists")
def test_reference_train_and_test_scripts_run_unchanged(tmp_path):
    from oracle import msunet_oracle as O
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys
    root = str(tmp_path)
    splits = _make_dataset(root)
    # synthetic SegFace checkpoint in the naming network/MSUNet.py:83-146 expects (oracle.encoder_checkpoint)
    m = MSUNetSys(img_size=S, embed_dim=32, depths=[2, 2, 2, 2], num_heads=[1, 2, 4, 8])
    ck, _ = O.encoder_checkpoint(m.state_dict(), "backbone.0.")
    g = torch.Generator().manual_seed(1)
    ck = {k: (torch.randn(v.shape, generator=g) * 0.02 + (1.0 if k.endswith("norm.weight") or "norm1.weight" in k or "norm2.weight" in k
                                                           else 0.0)) if v.is_floating_point() else v for k, v in ck.items()}
    ckpt = os.path.join(root, "segface.pt")
    torch.save({"state_dict_backbone": ck}, ckpt)
    out = os.path.join(root, "model_out")
    cfg = os.path.join(root, "cfg.yaml")
    with open(cfg, "w") as f:
        f.write(CFG.format(data=os.path.join(root, "data"), S=S, ckpt=ckpt, out=out, lists=os.path.join(root, "lists")))

    # ---------------- train.py (with the fused AdamW patched in, as INTEGRATION.md describes)
    rep = _run("train.py", os.path.join(root, "train_report.json"), ["--cfg", cfg], root, {"MSU_PATCH_ADAMW": "1"})
    pkg = "semantic_segmentation_of_stylegan2_artifacts_b200"
    assert rep["msunet"].startswith(pkg) and rep["loss"].startswith(pkg) and rep["metrics"].startswith(pkg)
    assert rep["csv_handler_file"].startswith(os.path.abspath(REF))          # the rest of `scripts` is still the reference's
    assert rep["adamw"].startswith(pkg) and rep["launches"] > 1000 and rep["lib"].endswith("libmsunet_sm100.so")
    best = os.path.join(out, "best_model.pth")
    assert os.path.exists(best) and os.path.exists(os.path.join(out, "epoch_2.pth"))
    payload = torch.load(best, map_location="cpu", weights_only=False)
    assert set(payload) == {"model", "epoch", "best_score"}
    assert list(payload["model"].keys()) == ["ms_unet." + k for k in m.state_dict().keys()]
    last = torch.load(os.path.join(out, "epoch_2.pth"), map_location="cpu", weights_only=False)
    st = last["optimizer"]["state"]
    assert len(st) > 200 and all(set(v) == {"step", "exp_avg", "exp_avg_sq"} for v in st.values())
    scal = json.load(open(os.path.join(out, "log", "scalars.json")))["info/total_loss"]
    losses = [v for _, v in scal]
    n_batches = (6 + 4) // 2
    assert len(losses) == 3 * n_batches and all(np.isfinite(losses))
    assert np.mean(losses[-n_batches:]) < 0.9 * np.mean(losses[:n_batches]), losses      # it learns
    import csv
    rows = list(csv.reader(open(os.path.join(out, "val_metric_all_epoch.csv"))))
    assert rows[0][:4] == ["epoch", "mean_accuracy", "mean_val_loss", "mean_train_loss"] and len(rows) == 4
    assert [r[0] for r in rows[1:]] == ["1", "2", "3"] and all(np.isfinite(float(r[6])) for r in rows[1:])
    assert len(list(csv.reader(open(os.path.join(out, "val_metric_fake_epoch.csv"))))) == 4
    preds = os.listdir(os.path.join(out, "final_preds"))
    assert any(p.endswith("_overlay_color.png") for p in preds) and any(p.endswith("_bin_mask.png") for p in preds)

    # ---------------- test.py: strict load of best_model.pth, metrics over the test split, heat-maps
    tout = os.path.join(root, "test_out")
    rep2 = _run("test.py", os.path.join(root, "test_report.json"), ["--cfg", cfg, "--check_point_dir", out, "--out_dir", tout], root)
    assert rep2["msunet"].startswith(pkg) and rep2["launches"] > 100
    run_dir = os.path.join(tout, os.listdir(tout)[0])
    rows = list(csv.reader(open(os.path.join(run_dir, "val_metric_all_epoch.csv"))))
    assert len(rows) == 2 and rows[1][0] == "1" and np.isfinite(float(rows[1][6]))
    pr = os.listdir(os.path.join(run_dir, "predictions"))
    for n in splits["test"]:
        assert f"{n}_bin_mask.png" in pr and f"{n}_heatmap.png" in pr and f"{n}_overlay_color.png" in pr
    log = open(os.path.join(run_dir, "log.txt")).read()
    assert "mean_dice_test" in log
