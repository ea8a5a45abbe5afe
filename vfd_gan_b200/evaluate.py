"""Device versions of the host detours ``MyGAN.test`` takes on every batch:
``threshold`` + ``morphology_proc`` (lib/utils.py:139-152) and the metrics ``lib/evaluate.py:14-91`` computes
with sklearn on the flattened voxel arrays. ``lib/evaluate.py`` itself keeps working unchanged on our outputs;
these functions give the same numbers without moving the voxel arrays to the host.
"""
import torch

from . import ops


def threshold(data, thr=0.5):
    """(data > 0.5).float() (lib/utils.py:149-152) -- returned together with the opening by ``threshold_open``."""
    return threshold_open(data, thr)[0]


def threshold_open(predict, thr=0.5):
    """predict fp32 (B,1,D,H,W) -> (t_pre, m_pre): the thresholded mask and its 5x5 opening, exactly as
    ``morphology_proc(threshold(predict))`` produces them (OpenCV treats each clip's (D,H,W) array as a D x H
    image with W channels, so the opening runs in the (D,H) plane)."""
    if not predict.is_cuda:
        raise RuntimeError("threshold_open: vfd_gan_b200 has no CPU path")
    if predict.dim() != 5 or predict.shape[1] != 1:
        raise RuntimeError(f"threshold_open expects (B,1,D,H,W), got {tuple(predict.shape)}")
    p = predict.contiguous().float()
    t = torch.empty_like(p)
    m = torch.empty_like(p)
    ops.threshold_open_op(p, float(thr), t, m)
    return t, m


def morphology_proc(video):
    """5x5 opening of an already binary (B,1,D,H,W) mask (lib/utils.py:139-147)."""
    return threshold_open(video, 0.5)[1]


def confusion_counts(labels, scores, thr, counts=None):
    """Accumulates [TP, FP, FN, TN] (int64 device tensor) with prediction = scores >= thr, label = labels > 0.5."""
    if not scores.is_cuda:
        raise RuntimeError("confusion_counts: vfd_gan_b200 has no CPU path")
    if counts is None:
        counts = torch.zeros(4, dtype=torch.int64, device=scores.device)
    ops.confusion_counts_op(labels.contiguous().float().view(-1), scores.contiguous().float().view(-1), float(thr),
                            counts)
    return counts


def binary_metrics_from_counts(tp, fp, fn, tn):
    """ROC area, PR area and F1 that lib/evaluate.py's ``roc`` / ``pr`` / ``f1_score`` branches return when the
    scores are a binary mask (the MyGAN.test case: ``m_pre_`` is a thresholded, opened mask):

    * ``roc_curve`` then has the operating points (0,0), (FPR,TPR), (1,1) -> area (1 + TPR - FPR) / 2;
    * ``precision_recall_curve`` has (recall, precision) = (1, P/n), (TPR, TP/(TP+FP)), (0, 1) and ``auc`` is the
      trapezoid rule over them;
    * ``f1_score`` of the mask binarised at 0.20 (a no-op on a {0,1} mask) = 2TP / (2TP + FP + FN)."""
    tp, fp, fn, tn = float(tp), float(fp), float(fn), float(tn)
    pos, neg = tp + fn, fp + tn
    tpr = tp / pos if pos else float("nan")
    fpr = fp / neg if neg else float("nan")
    roc = 0.5 * (1.0 + tpr - fpr)
    if tp + fp > 0:
        prec = tp / (tp + fp)
        base = pos / (pos + neg)
        # recall decreasing 1 -> tpr -> 0 with precision base -> prec -> 1; sklearn's auc integrates |dx|
        pr = 0.5 * (1.0 - tpr) * (base + prec) + 0.5 * tpr * (prec + 1.0)
    else:       # no predicted positives: the curve is (1, base) -> (0, 1)
        base = pos / (pos + neg) if pos + neg else float("nan")
        pr = 0.5 * (base + 1.0)
    f1 = 2.0 * tp / (2.0 * tp + fp + fn) if (2.0 * tp + fp + fn) > 0 else 0.0
    return {"roc": roc, "pr": pr, "f1": f1}


def roc_auc(labels, scores):
    """Exact ROC area of up to 16384 (score, label) pairs on the device (the per-clip anomaly-score sweep);
    equals ``lib/evaluate.py``'s ``roc`` = ``auc(*roc_curve(labels, scores)[:2])``. Returns a device double[3]:
    area, #positives, #negatives (no host synchronisation)."""
    if not scores.is_cuda:
        raise RuntimeError("roc_auc: vfd_gan_b200 has no CPU path")
    out = torch.empty(3, dtype=torch.float64, device=scores.device)
    ops.roc_auc_op(scores.contiguous().float().view(-1), labels.contiguous().float().view(-1), out)
    return out
