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


ROC_SINGLE_BLOCK_MAX = 16384


def roc_auc(labels, scores):
    """Exact ROC and precision-recall areas of (score, label) pairs on the device; equals ``lib/evaluate.py``'s
    ``roc`` = ``auc(*roc_curve(labels, scores)[:2])`` and ``pr`` = ``auc(recall, precision)``. Returns a device
    double[4]: ROC area, #positives, #negatives, PR area -- no host synchronisation.

    Up to 16384 pairs (the per-clip anomaly-score sweep) run in one block; anything larger (the voxel-level evaluation
    of test.py:175-202: every voxel of the test set) goes through the multi-block radix sort of ``csrc/roc_large.cu``
    (about 16 bytes of scratch per pair, n < 2^31)."""
    if not scores.is_cuda:
        raise RuntimeError("roc_auc: vfd_gan_b200 has no CPU path")
    out = torch.empty(4, dtype=torch.float64, device=scores.device)
    s, l = scores.contiguous().float().view(-1), labels.contiguous().float().view(-1)
    if s.numel() != l.numel():
        raise RuntimeError(f"roc_auc: {s.numel()} scores but {l.numel()} labels")
    if s.numel() <= ROC_SINGLE_BLOCK_MAX:
        ops.roc_auc_op(s, l, out)
    else:
        from . import _lib
        need = int(_lib.lib().vfd_roc_auc_large_workspace(s.numel()))
        ws = torch.empty(need, dtype=torch.uint8, device=s.device)
        ops.roc_auc_large_op(s, l, out, ws)
    return out


class VoxelCurveAccumulator:
    """The evaluation loop of test.py:175-202 without the host arrays: per batch ``add(gt, predict)`` keeps the
    flattened voxels of the raw prediction and of the ground truth on the device (the reference appends
    ``predict.permute(0,2,3,4,1).cpu().numpy()`` to a list and stacks it) and folds them into the confusion counts of
    the F1 branch (threshold 0.20, lib/evaluate.py:20-24); ``result()`` runs the device ROC / PR areas over all voxels
    seen and does the one device->host read. The order of the voxels does not enter any of the three metrics, so no
    permute is needed. ``capacity`` voxels are allocated up front (two fp32 arrays) and doubled on demand."""

    def __init__(self, device, capacity=1 << 24):
        self.scores = torch.empty(int(capacity), dtype=torch.float32, device=device)
        self.labels = torch.empty(int(capacity), dtype=torch.float32, device=device)
        self.counts = torch.zeros(4, dtype=torch.int64, device=device)
        self.n = 0

    def add(self, gt, predict):
        if not predict.is_cuda:
            raise RuntimeError("VoxelCurveAccumulator: vfd_gan_b200 has no CPU path")
        p, g = predict.detach().reshape(-1).float(), gt.detach().reshape(-1).float()
        if p.numel() != g.numel():
            raise RuntimeError(f"VoxelCurveAccumulator.add: {p.numel()} predictions but {g.numel()} labels")
        m = p.numel()
        if self.n + m > self.scores.numel():
            cap = max(2 * self.scores.numel(), self.n + m)
            for name in ("scores", "labels"):
                grown = torch.empty(cap, dtype=torch.float32, device=self.scores.device)
                grown[:self.n].copy_(getattr(self, name)[:self.n])
                setattr(self, name, grown)
        self.scores[self.n:self.n + m].copy_(p)
        self.labels[self.n:self.n + m].copy_(g)
        if m:
            confusion_counts(g, p, 0.20, self.counts)
        self.n += m

    def result(self):
        """-> {"roc", "pr", "f1"} as lib/evaluate.py's three metrics return them for (gts, predicts)."""
        if self.n == 0:
            raise RuntimeError("VoxelCurveAccumulator.result() before any add()")
        area = roc_auc(self.labels[:self.n], self.scores[:self.n]).tolist()
        tp, fp, fn, _ = self.counts.tolist()
        f1 = 2.0 * tp / (2.0 * tp + fp + fn) if (2.0 * tp + fp + fn) > 0 else 0.0
        return {"roc": area[0], "pr": area[3], "f1": f1}


TEST_KEYS = ("d/err_d_real_s/test", "d/err_d_real_t/test", "d/err_d_fake_s/test", "d/err_d_fake_t/test",
             "d/err_d_real/test", "d/err_d_fake/test", "d/err_d/test", "g/err_g_adv_s/test", "g/err_g_adv_t/test",
             "g/err_g_adv/test", "g/err_g_con/test", "g/err_g/test")


class GanEvaluator:
    """``MyGAN.test`` (models/mygannet.py:369-475) without its host detours: per batch the generator, threshold +
    5x5 opening, both optical flows, the discriminator on (gt, gt_flow) and (predict, pre_flow) and the twelve
    losses; the flattened (gt, m_pre) voxel arrays the reference hands to sklearn are reduced to confusion counts
    on the fly. ``result()`` does the one device->host read and returns the dictionaries the reference logs
    (``errors_dict`` means :459-472, ``score_dict`` :454-458).

    Like the reference, it does not switch the nets to ``eval()`` (BatchNorm keeps using batch statistics and
    updating its running ones) -- call ``.eval()`` yourself if that is not what you want -- and it
    keeps the reference's quirk that ``err_g`` is built from the temporal adversarial term only (:414)."""

    def __init__(self, netg, netd, w_adv=1, w_con=10, pos_weight=2):
        self.netg, self.netd = netg, netd
        self.w_adv, self.w_con, self.pos_weight = w_adv, w_con, pos_weight
        dev = next(netg.parameters()).device
        self.sums = torch.zeros(len(TEST_KEYS), dtype=torch.float64, device=dev)
        self.counts = torch.zeros(4, dtype=torch.int64, device=dev)
        self.batches = 0

    @torch.no_grad()
    def add_batch(self, inp, gt):
        """inp (B,3,D,H,W), gt (B,1,D,H,W) on the GPU -> (predict, t_pre, m_pre) of the batch (device tensors)."""
        import torch.nn.functional as F
        from .flow import video_to_flow
        netg, netd = self.netg, self.netd
        logits, _ = netg.forward_cl(ops.PackFn.apply(inp, 0))
        predict = ops.SigmoidHeadFn.apply(logits)
        t_pre, m_pre = threshold_open(predict)
        gt_flow = video_to_flow(gt.expand(-1, 3, -1, -1, -1))
        pre_flow = video_to_flow(predict.expand(-1, 3, -1, -1, -1))
        s_pr, s_fr, t_pr, t_fr = netd.forward_cl(ops.PackFn.apply(gt, 3), ops.PackFn.apply(gt_flow, 0))
        s_pf, s_ff, t_pf, t_ff = netd.forward_cl(ops.PackFn.apply(predict, 3), ops.PackFn.apply(pre_flow, 0))
        adv_s = ops.mse_cl(s_fr, s_ff, netd.spatdisc.feat_channels)
        adv_t = ops.mse_cl(t_fr, t_ff, netd.tempdisc.feat_channels)
        con = ops.WeightedBceFn.apply(predict, gt, float(self.pos_weight))
        ones, zeros = torch.ones_like(s_pr), torch.zeros_like(s_pf)
        e_rs, e_rt = F.binary_cross_entropy(s_pr, ones), F.binary_cross_entropy(t_pr, ones)
        e_fs, e_ft = F.binary_cross_entropy(s_pf, zeros), F.binary_cross_entropy(t_pf, zeros)
        real, fake = (e_rs + e_rt) * 0.5, (e_fs + e_ft) * 0.5
        self.sums += torch.stack([e_rs, e_rt, e_fs, e_ft, real, fake, (real + fake) * 0.5, adv_s, adv_t, adv_s + adv_t,
                                  con, adv_t * self.w_adv + con * self.w_con]).double()
        confusion_counts(gt, m_pre, 0.20, self.counts)
        self.batches += 1
        return predict, t_pre, m_pre

    def result(self):
        """-> (errors_dict, score_dict): per-batch means of the twelve losses; roc / pr / f1 of the opened masks."""
        vals = (self.sums / max(self.batches, 1)).tolist()
        scores = binary_metrics_from_counts(*self.counts.tolist())
        return dict(zip(TEST_KEYS, vals)), {"score/roc": scores["roc"], "score/pr": scores["pr"], "score/f1": scores["f1"]}
