"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, and the host modules keep the reference's names / signatures / state_dict keys."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

PKG = "semantic_segmentation_of_stylegan2_artifacts_b200"


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, PKG, "libmsunet_sm100.so")):
        g.build()
    import importlib
    return importlib.import_module(PKG)


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "msunet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(msu_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(built.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    from semantic_segmentation_of_stylegan2_artifacts_b200 import _lib
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert built.lib().msu_version() >= 100


def test_struct_layout_matches_header(built):
    from semantic_segmentation_of_stylegan2_artifacts_b200._lib import MsuEpilogue, MsuOperand
    assert ctypes.sizeof(MsuOperand) == built.lib().msu_struct_size(0) == 88
    assert ctypes.sizeof(MsuEpilogue) == built.lib().msu_struct_size(1) == 184      # 128 + the seven lnd_* pointers


@pytest.mark.parametrize("name,kw,img", [("t32_160", dict(embed_dim=32, depths=[2, 2, 2, 2], num_heads=[1, 2, 4, 8]), 160),
                                         ("t96_224", dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24]), 224)])
def test_state_dict_contract(built, name, kw, img):
    """Keys, order and shapes equal the reference's MSUNetSys.state_dict() (golden from the reference)."""
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = MSUNetSys(img_size=img, drop_path_rate=0.0, **kw)
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["sd_keys"])
    assert [",".join(map(str, v.shape)) for v in sd.values()] == list(g["sd_shapes"])
    idx = sd["layers.0.blocks.0.attn.relative_position_index"]
    from oracle import msunet_oracle as O
    assert idx.dtype == torch.int64 and torch.equal(idx, O.relative_position_index())


def test_signatures_match_reference(built):
    from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.MSUNet import MSUNet
    from semantic_segmentation_of_stylegan2_artifacts_b200.scripts import validation_functions as VF
    assert list(inspect.signature(MSUNet.__init__).parameters) == ["self", "config", "img_size", "num_classes", "zero_head", "vis"]
    p = inspect.signature(DynamicLoss.__init__).parameters
    assert [(k, v.default) for k, v in p.items()][1:] == [("roi_thresh", 0.04), ("alpha", 0.4), ("beta", 0.6), ("tversky_bce_mix", 0.5)]
    assert list(inspect.signature(VF.calculate_metrics).parameters) == [
        "model", "logging", "testloader", "dynamic_loss", "csv_all_epoch", "csv_fake_epoch", "csv_real_epoch",
        "csv_batch_real", "csv_batch_fake", "mean_train_loss", "epoch", "device", "split", "img_size", "sig_threshold",
        "output_num"]
    assert list(inspect.signature(VF.validation_loss).parameters) == ["model", "device", "val_loader", "dynamic_loss", "bool_break", "n_batches"]
    for fn in ("calculate_metrics_real", "calculate_metrics_fake"):
        assert list(inspect.signature(getattr(VF, fn)).parameters) == ["pred_bin", "pred", "ground_truth"]
    assert list(inspect.signature(VF.atrifact_prediction).parameters) == ["model", "testloader", "device", "img_size"]


def _reference_dir():
    for d in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(d, "network", "MSUNet.py")):
            return d
    return None


@pytest.mark.skipif(_reference_dir() is None, reason="no reference checkout (/root/reference or baseline/_ref)")
def test_signatures_match_the_reference_sources(built):
    """Where the reference is present its own sources are the authority: names, order and defaults of every public callable of
    the boundary are read from network/MSUNet.py, loss/DynamicLoss.py and scripts/validation_functions.py with `ast`
    (importing validation_functions needs medpy) and compared with this package's."""
    import ast
    from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.MSUNet import MSUNet
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys
    from semantic_segmentation_of_stylegan2_artifacts_b200.scripts import validation_functions as VF
    ref = _reference_dir()

    def ref_sig(rel, qual):
        tree = ast.parse(open(os.path.join(ref, rel)).read())
        scope = tree.body
        for part in qual.split("."):
            node = next(n for n in scope if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name == part)
            scope = getattr(node, "body", [])
        a = node.args
        names = [x.arg for x in a.args]
        defaults = [None] * (len(names) - len(a.defaults)) + [ast.literal_eval(d) if not isinstance(d, (ast.Attribute, ast.Name)) else "<expr>"
                                                                for d in a.defaults]
        return list(zip(names, defaults))

    def my_sig(fn):
        out = []
        for k, v in inspect.signature(fn).parameters.items():
            if v.kind in (v.VAR_KEYWORD, v.VAR_POSITIONAL):
                continue
            out.append((k, None if v.default is inspect._empty else v.default))
        return out

    def same(mine, theirs, extra_ok=()):
        mine = [m for m in mine if m[0] not in extra_ok]
        assert [m[0] for m in mine] == [t[0] for t in theirs], (mine, theirs)
        for (k, dm), (_, dt) in zip(mine, theirs):
            if dt != "<expr>":
                assert dm == dt or (dm is None and dt is None), (k, dm, dt)

    same(my_sig(MSUNet.__init__), ref_sig("network/MSUNet.py", "MSUNet.__init__"))
    same(my_sig(MSUNet.forward), ref_sig("network/MSUNet.py", "MSUNet.forward"))
    for meth in ("freeze_encoder", "unfreeze_encoder", "load_segface_weight", "load_IMAGENET1K_weight"):
        same(my_sig(getattr(MSUNet, meth)), ref_sig("network/MSUNet.py", "MSUNet." + meth))
    same(my_sig(MSUNetSys.__init__), ref_sig("network/model_parts.py", "MSUNetSys.__init__"), extra_ok=("run_dead_branches",))
    same(my_sig(DynamicLoss.__init__), ref_sig("loss/DynamicLoss.py", "DynamicLoss.__init__"))
    same(my_sig(DynamicLoss.forward), ref_sig("loss/DynamicLoss.py", "DynamicLoss.forward"))
    for fn in ("validation_loss", "calculate_metrics", "calculate_metrics_real", "calculate_metrics_fake", "atrifact_prediction"):
        same(my_sig(getattr(VF, fn)), ref_sig("scripts/validation_functions.py", fn))


def test_errors_and_no_cpu_fallback(built):
    from types import SimpleNamespace as NS
    from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.MSUNet import MSUNet
    cfg = NS(MODEL=NS(SWIN=NS(PATCH_SIZE=4, IN_CHANS=3, EMBED_DIM=32, DEPTHS=[2, 2, 2, 2], NUM_HEADS=[1, 2, 4, 8],
                              WINDOW_SIZE=7, MLP_RATIO=4.0, QKV_BIAS=True, APE=False, PATCH_NORM=True),
                      DROP_RATE=0.0, DROP_PATH_RATE=0.1, ATTN_DROP_RATE=0.0),
             TRAIN=NS(USE_CHECKPOINT=False))
    m = MSUNet(cfg, img_size=64, num_classes=1)
    assert next(iter(m.state_dict())).startswith("ms_unet.patch_embed.")
    with pytest.raises(ValueError):
        m(torch.zeros(1, 4, 64, 64))                      # channel check, network/MSUNet.py:48-51
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64))                      # CPU tensors: loud failure, never a fallback
    with pytest.raises(ValueError):
        DynamicLoss()(torch.zeros(2, 1, 8, 8), torch.zeros(3, 8, 8))   # loss/DynamicLoss.py:93-94
    with pytest.raises(RuntimeError):
        DynamicLoss()(torch.zeros(2, 1, 8, 8), torch.zeros(2, 8, 8))
    from semantic_segmentation_of_stylegan2_artifacts_b200 import ops
    from semantic_segmentation_of_stylegan2_artifacts_b200.data import CudaPrefetcher
    with pytest.raises(RuntimeError):
        ops.stage_u8(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))          # input staging: CUDA only as well
    with pytest.raises(ValueError):
        ops.stage_u8(torch.zeros(1, 8, 8, 3))                             # dataset/dataset.py:41 hands over uint8
    with pytest.raises(RuntimeError):
        CudaPrefetcher([], "cpu")
    m.freeze_encoder(True)
    assert not any(p.requires_grad for p in m.ms_unet.layers.parameters())
    m.unfreeze_encoder(0)
    assert all(p.requires_grad for p in m.ms_unet.patch_embed.parameters())
    with pytest.raises(ValueError):
        m.unfreeze_encoder(7)


def test_encoder_key_remap():
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.MSUNet import _remap
    assert _remap("backbone.0.0.0.weight", "backbone.0.") == "patch_embed.proj.weight"
    assert _remap("backbone.0.0.2.bias", "backbone.0.") == "patch_embed.norm.bias"
    assert _remap("backbone.0.1.1.attn.qkv.weight", "backbone.0.") == "layers.0.blocks.1.attn.qkv.weight"
    assert _remap("backbone.0.5.17.mlp.3.bias", "backbone.0.") == "layers.2.blocks.17.mlp.3.bias"
    assert _remap("features.6.reduction.weight", "features.") == "layers.2.downsample.reduction.weight"
    assert _remap("features.7.0.norm1.weight", "features.") == "layers.3.blocks.0.norm1.weight"
    assert _remap("features.9.x", "features.") is None


def test_dropin_overlay_resolves_reference_imports(built):
    """With dropin/ ahead on PYTHONPATH the reference's own import lines (train.py:10, trainer.py:27,29) bind to this repo."""
    import subprocess
    import sys
    code = ("from network.MSUNet import MSUNet; from loss.DynamicLoss import DynamicLoss; "
            "from scripts.validation_functions import calculate_metrics, validation_loss; "
            "import network.model_parts as mp; "
            "print(MSUNet.__module__, DynamicLoss.__module__, calculate_metrics.__module__, mp.MSUNetSys.__module__)")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, PKG, "dropin"), ROOT]))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert all(m.startswith(PKG + ".") for m in out.stdout.split()), out.stdout


def test_checkpoint_roundtrip_legacy_format(built, tmp_path):
    """best_model.pth interchange (trainer.py:372-379 saves, test.py:97-109 loads with strict=True): a state_dict written in
    the legacy (non-zip) serialisation and in the default one loads back bit-identically, keys in the reference's order."""
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys
    kw = dict(img_size=64, embed_dim=32, depths=[2, 2, 2, 2], num_heads=[1, 2, 4, 8])
    m = MSUNetSys(**kw)
    sd = m.state_dict()
    for legacy in (True, False):
        f = tmp_path / f"best_model_{legacy}.pth"
        torch.save({"epoch": 3, "model": sd}, f, _use_new_zipfile_serialization=not legacy)
        ck = torch.load(f, map_location="cpu", weights_only=False)
        m2 = MSUNetSys(**kw)
        missing = m2.load_state_dict(ck["model"], strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        assert list(m2.state_dict().keys()) == list(sd.keys())
        assert all(torch.equal(a, b) for a, b in zip(m2.state_dict().values(), sd.values()))


@pytest.mark.parametrize("what,root,wrap", [("segface", "backbone.0.", "state_dict_backbone"), ("imagenet", "features.", None)])
def test_pretrained_encoder_interchange_matches_reference(built, tmp_path, what, root, wrap):
    """load_segface_weight / load_IMAGENET1K_weight (network/MSUNet.py:63-240) put every checkpoint tensor where the
    reference's loaders put it (fixture recorded by running them: oracle/make_remap_golden.py), leave the decoder alone, skip the
    SegFace head, and fail like the reference on unknown keys, mismatching shapes and a missing wrapper key."""
    import json
    import logging
    from types import SimpleNamespace as NS
    from oracle import msunet_oracle as O
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.MSUNet import MSUNet
    g = json.load(open(os.path.join(GOLDEN, "encoder_remap.json")))
    path = str(tmp_path / (what + ".pth"))

    def cfg():
        return NS(MODEL=NS(SWIN=NS(PATCH_SIZE=4, IN_CHANS=3, EMBED_DIM=g["embed_dim"], DEPTHS=g["depths"], NUM_HEADS=g["num_heads"],
                                   WINDOW_SIZE=7, MLP_RATIO=4.0, QKV_BIAS=True, APE=False, PATCH_NORM=True),
                           DROP_RATE=0.0, DROP_PATH_RATE=0.1, ATTN_DROP_RATE=0.0, PRETRAIN_SEGFACE=path, PRETRAIN_IMAGENET1K=path),
                  TRAIN=NS(USE_CHECKPOINT=False))

    def fresh():
        m = MSUNet(cfg(), img_size=g["img_size"], num_classes=1)
        with torch.no_grad():
            for v in m.ms_unet.state_dict().values():
                v.fill_(-1)
        return m

    def load(m):
        (m.load_segface_weight if what == "segface" else m.load_IMAGENET1K_weight)(cfg(), logging)

    m = fresh()
    ckpt, where = O.encoder_checkpoint(m.ms_unet.state_dict(), root)
    extra = "backbone.1.classifier.weight" if what == "segface" else "head.weight"       # SegFace head / classifier: ignored
    ckpt[extra] = torch.zeros(3)
    by_id = {int(v.flatten()[0]): k for k, v in ckpt.items() if k.startswith(root)}
    torch.save({wrap: ckpt} if wrap else ckpt, path)
    load(m)
    got = {k: (by_id[int(v.flatten()[0])] if float(v.flatten()[0]) > 0 else None) for k, v in m.ms_unet.state_dict().items()}
    assert got == g[what]
    assert sum(v is not None for v in got.values()) == len(where) and all(got[k] == nk for k, nk in where.items())
    # error behaviour of the reference
    bad = dict(ckpt)
    bad[root + "9.0.norm1.weight"] = torch.zeros(3)                                        # no such stage
    torch.save({wrap: bad} if wrap else bad, path)
    with pytest.raises(ValueError):
        load(fresh())
    bad = dict(ckpt)
    k0 = next(k for k in bad if k.endswith("qkv.weight"))
    bad[k0] = torch.zeros(5, 5)
    torch.save({wrap: bad} if wrap else bad, path)
    with pytest.raises(ValueError):
        load(fresh())
    if wrap:
        torch.save(ckpt, path)                                                             # wrapper key missing
        with pytest.raises(KeyError):
            load(fresh())
    torch.save({wrap: {"other.weight": torch.zeros(1)}} if wrap else {"other.weight": torch.zeros(1)}, path)
    with pytest.raises(ValueError):                                                        # "No new keys from backbone!!"
        load(fresh())


def test_host_helpers_refuse_cpu(built):
    """optim.FusedAdamW and graphs.GraphedStep are CUDA-only like the hot path itself."""
    from semantic_segmentation_of_stylegan2_artifacts_b200.graphs import GraphedStep
    from semantic_segmentation_of_stylegan2_artifacts_b200.optim import FusedAdamW, patch_torch_adamw
    p = torch.zeros(3, requires_grad=True)
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError):
        FusedAdamW([p]).step()
    with pytest.raises(NotImplementedError):
        FusedAdamW([p], amsgrad=True)
    with pytest.raises(ValueError):
        FusedAdamW([p], lr=-1.0)
    with pytest.raises(RuntimeError):
        GraphedStep(torch.nn.Linear(2, 2), lambda a, b: a.sum(), torch.zeros(1, 2), torch.zeros(1))
    keep = torch.optim.AdamW
    try:
        patch_torch_adamw()
        assert torch.optim.AdamW is FusedAdamW
    finally:
        torch.optim.AdamW = keep


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours): exactly one stdout line, the contract's keys."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    # the reference's own modules when baseline/_ref holds them (tools/install_reference.py), else the oracle port
    want_kind = "reference" if os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "network", "model_parts.py")) else "port"
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    ours = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_n1.json")))
    assert (d["metric"], d["unit"]) == (ours["metric"], ours["unit"])
