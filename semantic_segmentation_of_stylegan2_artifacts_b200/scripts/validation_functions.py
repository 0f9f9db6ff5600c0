"""Validation / metric entry points with the reference's signatures and return values
(scripts/validation_functions.py:8-357).  All per-pixel work — sigmoid, threshold, TP/FP/FN/TN
counting and the eight soft sums — is one fused CUDA pass per call (msu_metrics); the derived
ratios, the epoch aggregation, CSV rows and the Score stay host-side Python as in the reference.
medpy is not needed: dc/jc/precision/recall are evaluated from the exact integer counts.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops

try:  # progress bars are optional plumbing
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(it, **kw):
        return it


def _as_u8(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    return t.view(torch.uint8) if t.dtype == torch.bool else (t != 0).view(torch.uint8)


def _count(pred_bin, pred, ground_truth):
    """One fused pass -> (tp, fp, fn, tn) python ints and the 8 soft sums as python floats."""
    if not pred.is_cuda:
        raise RuntimeError("metric counting runs on CUDA only (no CPU fallback); move the tensors to the GPU")
    p = pred.contiguous()
    if p.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        p = p.float()
    counts, soft, _ = ops.metrics(p.view(1, -1), _as_u8(ground_truth).view(1, -1), _as_u8(pred_bin).view(1, -1),
                                  from_logits=False, thr=0.0)
    c = counts.cpu().tolist()[0]
    s = soft.cpu().tolist()[0]
    return c, s


def image_counts_from_logits(out_logits: torch.Tensor, label: torch.Tensor, sig_threshold: float = 0.5,
                             want_pred: bool = True, prob_f32: bool = True):
    """Batched fused path: logits [B,1,H,W] + label [B,H,W] -> (counts int64 [B,4], soft float64 [B,8], pred).
    pred = sigmoid(logits) in fp32 whatever the logits dtype, pred_bin = pred > thr, gt = label > 0
    (scripts/validation_functions.py:106-108) — all inside one kernel, device-resident results.  The reference evaluates under
    fp16 autocast (:78): its sigmoid has 11 bits below 0.5's ulp; bf16 logits rounded to a bf16 sigmoid would move the decision
    boundary (ulp 2^-9 above 0.5) and shift the soft Dice / IoU that feed `Score`, so probabilities stay fp32 here
    (`prob_f32=False`: torch.sigmoid semantics on the logits dtype, i.e. rounded to it before the threshold)."""
    B = out_logits.shape[0]
    lg = out_logits.contiguous().view(B, -1)
    lb = label.contiguous().float().view(B, -1)
    counts, soft, pred = ops.metrics(lg, lb, None, from_logits=True, thr=float(sig_threshold), want_pred=want_pred, prob_f32=prob_f32)
    if pred is not None:
        pred = pred.view(B, *out_logits.shape[2:])
    return counts, soft, pred


# -------------------- Validation Loss ---------------------------- #
def validation_loss(model, device, val_loader, dynamic_loss, bool_break=False, n_batches=0):
    val_losses = []
    model.eval()
    with torch.inference_mode():
        for i_batch, sampled_batch in enumerate(val_loader):
            if bool_break and (i_batch >= n_batches):
                break
            image = sampled_batch["image"].to(device, non_blocking=True)
            label = sampled_batch["label"].to(device, non_blocking=True)
            assert image.ndim == 4
            assert label.ndim in (3, 4)
            out_logits = model(image)
            val_losses.append(dynamic_loss(out_logits, label))  # stays on device: one sync at the end
    model.train()
    if len(val_losses) <= 0:
        return float("nan")
    vals = torch.stack([v.float().reshape(()) for v in val_losses]).cpu().tolist()
    return sum(vals) / len(vals)


def _real_from_counts(c, s):
    tp, fp, fn, tn = c
    confusion_matrix_soft = [[float(s[0]), float(s[1])], [float(s[2]), float(s[3])]]
    FPR = fp / (fp + tn)
    total = tp + tn + fp + fn
    if total <= 0:
        raise ValueError(f"Real metric calculation failed because total = {total}")
    return [[tp, fp], [fn, tn]], confusion_matrix_soft, float((tp + tn) / total), FPR


def _fake_from_counts(c, s):
    smooth = 1e-8
    tp, fp, fn, tn = c
    n_pred, n_gt = tp + fp, tp + fn
    bin_dice = 2.0 * tp / float(n_pred + n_gt) if (n_pred + n_gt) else 0.0       # medpy dc
    bin_recall = tp / float(tp + fn) if (tp + fn) else 0.0                       # medpy recall
    bin_precision = tp / float(tp + fp) if (tp + fp) else 0.0                    # medpy precision
    bin_IoU = float(tp) / float(tp + fp + fn)                                    # medpy jc (no zero guard)
    bin_f1 = 2 * (bin_precision * bin_recall) / (bin_precision + bin_recall + smooth)
    total = tp + tn + fp + fn
    if total <= 0:
        raise ValueError(f"Real metric calculation failed because total = {total}")
    bin_accuracy = (tp + tn) / total
    confusion_matrix_soft = [[float(s[0]), float(s[1])], [float(s[2]), float(s[3])]]
    inter, sum_p_2, sum_g_2, sum_p, sum_g = s[0], s[4], s[5], s[6], s[7]
    i_soft_dice = float((2.0 * inter + smooth) / (sum_p_2 + sum_g_2 + smooth))
    i_soft_iou = float((inter + smooth) / (sum_p + sum_g - inter + smooth))
    return (float(bin_accuracy), float(bin_recall), float(bin_precision), float(bin_IoU), float(bin_dice),
            float(bin_f1), [[tp, fp], [fn, tn]], confusion_matrix_soft, i_soft_dice, i_soft_iou)


# -------------------- Calculating REAL Metrics ---------------------------- #
def calculate_metrics_real(pred_bin, pred, ground_truth):
    c, s = _count(pred_bin, pred, ground_truth)
    return _real_from_counts(c, s)


# -------------------- Calculating FAKEs Metrics ---------------------------- #
def calculate_metrics_fake(pred_bin, pred, ground_truth):
    c, s = _count(pred_bin, pred, ground_truth)
    return _fake_from_counts(c, s)


# -------------------- Metrics ----------------------------------------------- #
def calculate_metrics(model, logging, testloader, dynamic_loss, csv_all_epoch, csv_fake_epoch, csv_real_epoch,
                      csv_batch_real, csv_batch_fake, mean_train_loss, epoch, device=None, split="test", img_size=1024,
                      sig_threshold=0.5, output_num=10):
    if output_num >= len(testloader):
        output_num = len(testloader)
    patch_size = (img_size, img_size)
    model.eval()
    num_cases = 0
    output_saver = []
    real_conf_matrix_bin_list, real_confusion_matrix_soft_list, accuracy_list_real, FRP_list = [], [], [], []
    real_image_counter = 0
    confusion_matrix_soft_list, fake_conf_matrix_bin_list, fake_confusion_matrix_soft_list = [], [], []
    accuracy_list_fake, metric_fake_list, accuracy_list, confusion_matrix_bin_list = [], [], [], []

    # Device pass: every batch (any batch size — the reference asserts 1, validation_functions.py:89) is evaluated without a host
    # synchronisation; losses, integer counts, soft sums and the first `output_num` heat-maps stay on the device until the end.
    dev_loss, dev_counts, dev_soft, names, kept_pred = [], [], [], [], []
    with torch.inference_mode():
        for i_batch, sampled_batch in tqdm(enumerate(testloader), total=len(testloader)):
            image = sampled_batch["image"].to(device, non_blocking=True)
            loss_label = sampled_batch["label"].to(device, non_blocking=True)
            assert image.ndim == 4
            assert loss_label.ndim in (3, 4)
            B, C, H, W = image.shape
            assert ((H, W) != tuple(patch_size)) == False  # noqa: E712
            image = image.float()
            label = loss_label.squeeze(1) if loss_label.ndim == 4 else loss_label

            out_logits = model(image)
            if out_logits.shape[1] != 1:
                raise ValueError(f"Binary task expected 1 logit channel, got {out_logits.shape[1]}")
            if hasattr(dynamic_loss, "per_sample"):
                loss_b = dynamic_loss.per_sample(out_logits, loss_label)
            else:   # a foreign loss object: one call per image, as the reference does
                loss_b = torch.stack([dynamic_loss(out_logits[b:b + 1], loss_label[b:b + 1]).float().reshape(()) for b in range(B)])
            want_pred = len(kept_pred) < output_num
            counts, soft, pred = image_counts_from_logits(out_logits, label, sig_threshold, want_pred=want_pred)
            dev_loss.append(loss_b.float())
            dev_counts.append(counts)
            dev_soft.append(soft)
            case_names = sampled_batch['case_name']
            for b in range(B):
                names.append(case_names[b] if b < len(case_names) else f"{case_names[0]}_{b}")
                if len(kept_pred) < output_num and pred is not None:
                    kept_pred.append((names[-1], pred[b]))
    # one device -> host hop for the whole split
    all_loss = torch.cat(dev_loss).cpu().tolist() if dev_loss else []
    all_counts = torch.cat(dev_counts).cpu().tolist() if dev_counts else []
    all_soft = torch.cat(dev_soft).cpu().tolist() if dev_soft else []
    output_saver = [(n, p.detach().cpu()) for n, p in kept_pred]

    for val_loss, c, s in zip(all_loss, all_counts, all_soft):
        has_artifact = (c[0] + c[2]) > 0  # ground_truth.any()  (tp + fn = |gt|)
        if not has_artifact:
            real_image_counter += 1
            confusion_matrix_bin, confusion_matrix_soft, accuracy, FRP = _real_from_counts(c, s)
            confusion_matrix_bin_list.append(confusion_matrix_bin)
            real_conf_matrix_bin_list.append(confusion_matrix_bin)
            real_confusion_matrix_soft_list.append(confusion_matrix_soft)
            confusion_matrix_soft_list.append(confusion_matrix_soft)
            accuracy_list.append((accuracy, float(val_loss)))
            accuracy_list_real.append((accuracy, float(val_loss)))
            FRP_list.append(float(FRP))
        else:
            (bin_accuracy, bin_recall, bin_precision, bin_IoU, bin_dice, bin_f1, confusion_matrix_bin,
             confusion_matrix_soft, i_soft_dice, i_soft_iou) = _fake_from_counts(c, s)
            metric_fake_list.append([bin_accuracy, bin_recall, bin_precision, bin_IoU, bin_dice, bin_f1,
                                     i_soft_dice, i_soft_iou])
            confusion_matrix_bin_list.append(confusion_matrix_bin)
            fake_conf_matrix_bin_list.append(confusion_matrix_bin)
            confusion_matrix_soft_list.append(confusion_matrix_soft)
            fake_confusion_matrix_soft_list.append(confusion_matrix_soft)
            accuracy_list.append((bin_accuracy, float(val_loss)))
            accuracy_list_fake.append((bin_accuracy, float(val_loss)))
        num_cases += 1

    if num_cases == 0:
        logging.error(f"No {split} cases processed. Check your dataset/split.")
        raise ValueError(f"Expected at least one {split} cases")
    if len(metric_fake_list) == 0:
        raise ValueError(f"No valid fake {split} metrics to aggregate.")

    if real_image_counter > 0:
        mean_acc_and_loss = np.mean(np.array(accuracy_list_real, dtype=float), axis=0)
        mean_confusion_matrix_bin_real = np.mean(np.array(real_conf_matrix_bin_list, dtype=float), axis=0).flatten().tolist()
        mean_confusion_matrix_soft_real = np.mean(np.array(real_confusion_matrix_soft_list, dtype=float), axis=0).flatten().tolist()
        mean_FPR = np.mean(np.array(FRP_list, dtype=float), axis=0)
        (mean_accuracy_real, mean_val_loss_real) = mean_acc_and_loss
        csv_real_epoch.writerow([epoch, float(mean_accuracy_real), mean_confusion_matrix_bin_real,
                                 mean_confusion_matrix_soft_real, mean_val_loss_real, mean_FPR])
        logging.info(f"{split} real performance for epoch {epoch} :"
                     f" mean_confusion_matrix_bin [[tp, fp],[fn, tn]] {mean_confusion_matrix_bin_real} "
                     f" mean_accuracy {mean_accuracy_real} mean_val_loss{mean_val_loss_real}")

    (mean_accuracy_fake, mean_val_loss_fake) = np.mean(np.array(accuracy_list_fake, dtype=float), axis=0)
    mean_confusion_matrix_bin_fake = np.mean(np.array(fake_conf_matrix_bin_list, dtype=float), axis=0).flatten().tolist()
    mean_confusion_matrix_soft_fake = np.mean(np.array(fake_confusion_matrix_soft_list, dtype=float), axis=0).flatten().tolist()
    mean_fake_metric = np.mean(np.array(metric_fake_list, dtype=float), axis=0)
    (mean_bin_accuracy, mean_bin_recall, mean_bin_precision, mean_bin_IoU, mean_bin_dice, mean_bin_f1, mean_soft_dice,
     mean_soft_iou) = mean_fake_metric
    # as in the reference (:180) this raises NameError when the split holds no real image
    Score = mean_soft_dice - (10 * mean_FPR)
    csv_fake_epoch.writerow([epoch, float(mean_accuracy_fake), float(mean_val_loss_fake), mean_confusion_matrix_bin_fake,
                             mean_confusion_matrix_soft_fake, *[float(x) for x in mean_fake_metric]])
    logging.info(f"{epoch}_fake: mean_soft_dice {mean_soft_dice} mean_val_loss {mean_val_loss_fake} mean_bin_recall "
                 f"{mean_bin_recall} mean_bin_precision {mean_bin_precision} mean_bin_dice {mean_bin_dice}")
    (mean_accuracy, mean_val_loss) = np.mean(np.array(accuracy_list, dtype=float), axis=0)
    mean_confusion_matrix_bin = np.mean(np.array(confusion_matrix_bin_list, dtype=float), axis=0).flatten().tolist()
    mean_confusion_matrix_soft = np.mean(np.array(confusion_matrix_soft_list, dtype=float), axis=0).flatten().tolist()
    csv_all_epoch.writerow([epoch, float(mean_accuracy), float(mean_val_loss), float(mean_train_loss),
                            mean_confusion_matrix_bin, mean_confusion_matrix_soft, Score])
    logging.info(f"{split} epoch {epoch}: mean_accuracy {mean_accuracy} "
                 f"mean_cofusion_matrix [[tp, fp],[fn, tn]]{mean_confusion_matrix_bin} mean_val_loss {mean_val_loss}")
    print(f"epoch{epoch} val_loss:{mean_val_loss} train_loss:{mean_train_loss} mean_soft_dice:{mean_soft_dice} "
          f"mean_FRP {mean_FPR} Score {Score}")
    return mean_soft_dice, output_saver, Score, mean_FPR


def atrifact_prediction(model, testloader, device=None, img_size=1024):
    output_num = len(testloader)
    patch_size = (img_size, img_size)
    model.eval()
    output_saver = []
    with torch.inference_mode():
        for i_batch, sampled_batch in tqdm(enumerate(testloader), total=len(testloader)):
            image = sampled_batch["image"].to(device, non_blocking=True)
            case_name = sampled_batch['case_name'][0]
            assert image.ndim == 4
            assert image.shape[0] == 1
            B, C, H, W = image.shape
            assert ((H, W) != tuple(patch_size)) == False  # noqa: E712
            out_logits = model(image.float())
            if out_logits.shape[1] != 1:
                raise ValueError(f"Binary task expected 1 logit channel, got {out_logits.shape[1]}")
            dummy = torch.zeros(1, H, W, dtype=torch.float32, device=out_logits.device)
            _, _, pred = image_counts_from_logits(out_logits, dummy, 0.5, want_pred=True)
            if i_batch < output_num:
                output_saver.append((case_name, pred[0].detach().cpu()))
    return output_saver
