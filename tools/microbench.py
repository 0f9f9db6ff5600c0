"""BASELINE.json configs 4 and 5: device-timed micro-benchmarks, one JSON line each.

  attn   : shifted-window attention block (A6: LN1 output -> qkv GEMM -> attention core -> proj GEMM) forward and
           forward+backward, H in {56,64,128,256}, C in {96,192,384}, shift in {0,3}, batch sweep; reports ms,
           TFLOP/s (SURVEY §8d attention FLOPs) and the fraction of the measured bf16 peak.
  loss   : fused BCE+Tversky DynamicLoss forward+backward, S in {224,512,1024}, batch sweep; GB/s against the
           algorithmic bytes (fwd (e+4) B S^2, bwd (2e+4) B S^2) and the fraction of the measured HBM peak.
  infer  : config 4 — MS-UNet T96 inference at 1024x1024, batch 8, eval mode + fused sigmoid/threshold/TP-FP-FN-TN
           counting (img/s), counts checked bit-exactly against the reference formulas on the same logits.

Usage: python tools/microbench.py [attn|loss|stage|infer|all] [--quick]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import T96, synth_batch  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200 import ops  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.functional import window_geo  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.scripts.validation_functions import image_counts_from_logits  # noqa: E402

dev = torch.device("cuda:0")
bf = torch.bfloat16
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def attn_case(Bn, H, C, shift):
    nH = C // 32
    geo = window_geo(H, H, shift)
    nW = Bn * (geo[2] // 7) * (geo[3] // 7)
    Tw, T = nW * 49, Bn * H * H
    xw = torch.randn(Tw, C, device=dev).to(bf)            # LN1 output in window order
    x = torch.randn(T, C, device=dev).to(bf)              # residual stream
    wq = (torch.randn(3 * C, C, device=dev) * C ** -0.5).to(bf)
    wp = (torch.randn(C, C, device=dev) * C ** -0.5).to(bf)
    bq, bp = torch.randn(3 * C, device=dev) * 0.1, torch.randn(C, device=dev) * 0.1
    bias = ops.relbias_expand(torch.randn(169, nH, device=dev) * 0.1, nH)
    qkv = torch.empty(Tw, 3 * C, dtype=bf, device=dev)
    x1 = torch.empty(T, C, dtype=bf, device=dev)
    do = torch.randn(Tw, C, device=dev).to(bf)
    st = {}

    def fwd():
        ops.gemm(ops.operand(xw), ops.operand(wq), ops.epilogue(qkv, bias=bq), Tw, 3 * C, C, dev)
        st["o"] = ops.winattn_fwd(qkv, bias, nW, nH, geo)
        ops.gemm(ops.operand(st["o"]), ops.operand(wp), ops.epilogue(x1, bias=bp, R=x, map=ops.MAP_WINDOW, geo=geo), Tw, C, C, dev)

    def core_bwd():
        ops.winattn_bwd(qkv, bias, st["o"], do, nW, nH, geo)

    t_f = timed(fwd)
    t_b = timed(core_bwd)
    flops = 2 * Tw * C * 3 * C + 2 * (2 * nW * nH * 49 * 49 * 32) + 2 * Tw * C * C
    return {"bench": "window_attention", "B": Bn, "H": H, "C": C, "shift": shift, "windows": nW, "heads": nH,
            "fwd_ms": round(t_f, 4), "fwd_tflops": round(flops / t_f / 1e9, 1),
            "fwd_frac_of_measured_bf16_peak": round(flops / t_f / 1e9 / PEAKS["bf16_tflops"], 4),
            "core_bwd_ms": round(t_b, 4)}


def loss_case(Bn, S):
    crit = DynamicLoss(alpha=0.2, beta=0.8, tversky_bce_mix=0.45)
    logits = torch.randn(Bn, 1, S, S, device=dev).to(bf).requires_grad_(True)
    _, y = synth_batch(Bn, S, 7)
    y = y.to(dev)

    def step():
        logits.grad = None
        crit(logits, y).backward()

    t = timed(step)
    byts = (2 + 4) * Bn * S * S + (2 * 2 + 4) * Bn * S * S
    return {"bench": "dynamic_loss_fwd_bwd", "B": Bn, "S": S, "ms": round(t, 4), "GBps": round(byts / t / 1e6, 1),
            "frac_of_measured_hbm_peak": round(byts / t / 1e6 / PEAKS["hbm_gbs"], 4)}


def stage_case(Bn, S):
    """uint8 HWC batch -> fp32 CHW / 255 + flip + label > 127: 4 B read + 16 B written per pixel."""
    img = torch.randint(0, 256, (Bn, S, S, 3), dtype=torch.uint8, device=dev)
    lab = torch.randint(0, 256, (Bn, S, S), dtype=torch.uint8, device=dev)
    flip = (torch.arange(Bn, device=dev) % 2).to(torch.uint8)
    t = timed(lambda: ops.stage_u8(img, lab, flip))
    byts = 20 * Bn * S * S
    return {"bench": "stage_u8", "B": Bn, "S": S, "ms": round(t, 4), "GBps": round(byts / t / 1e6, 1),
            "frac_of_measured_hbm_peak": round(byts / t / 1e6 / PEAKS["hbm_gbs"], 4)}


def infer_case(Bn=8, S=1024, steps=5):
    torch.manual_seed(1234)
    m = MSUNetSys(img_size=S, drop_path_rate=0.1, **T96).to(dev).eval()
    x, y = synth_batch(Bn, S, 4321)
    x, y = x.to(dev), y.to(dev)
    with torch.inference_mode():
        for _ in range(2):
            logits = m(x)
            counts, soft, _ = image_counts_from_logits(logits, y, 0.5, want_pred=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            logits = m(x)
            counts, soft, _ = image_counts_from_logits(logits, y, 0.5, want_pred=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        # reference formulas on the same logits (scripts/validation_functions.py:106-108, 219-227)
        pred = torch.sigmoid(logits.squeeze(1).float())  # fp32 probabilities (msu_metrics thresholds in fp32 whatever the logits dtype)
        pb, gt = pred > 0.5, y > 0
        ref = torch.stack([(pb & gt).sum((1, 2)), (pb & ~gt).sum((1, 2)), (~pb & gt).sum((1, 2)), (~pb & ~gt).sum((1, 2))], 1)
        exact = bool((ref.cpu() == counts.cpu()).all())
    return {"bench": "inference_plus_dice_iou_counts", "B": Bn, "S": S, "ms_per_batch": round(ms, 3),
            "img_per_s": round(Bn / ms * 1e3, 2), "model_tflops": round(Bn / ms * 1e3 * 782.0 / 1e3, 1),
            "counts_bit_exact_vs_reference_formulas": exact}


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    quick = "--quick" in sys.argv
    if what in ("attn", "all"):
        Hs = [(56, 96), (64, 192), (128, 96), (256, 96), (64, 384)] if not quick else [(128, 96), (64, 192)]
        for H, C in Hs:
            for shift in (0, 3):
                for Bn in ((1, 16, 64) if not quick else (16,)):
                    if Bn * H * H * C > 64 * 128 * 128 * 96 * 2:
                        continue
                    print(json.dumps(attn_case(Bn, H, C, shift)), flush=True)
    if what in ("loss", "all"):
        for S in (224, 512, 1024):
            for Bn in ((1, 16, 64) if not quick else (16,)):
                print(json.dumps(loss_case(Bn, S)), flush=True)
    if what in ("stage", "all"):
        for Bn, S in ((16, 512), (64, 512), (16, 1024)):
            print(json.dumps(stage_case(Bn, S)), flush=True)
    if what in ("infer", "all"):
        print(json.dumps(infer_case()), flush=True)


if __name__ == "__main__":
    main()
